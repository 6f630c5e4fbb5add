"""RACE fine-tuning (parity: tasks/race/finetune.py)."""
from megatron_llm_b200 import get_args, get_tokenizer, print_rank_0
from megatron_llm_b200.models.enums import ModelType
from megatron_llm_b200.models.multiple_choice import MultipleChoice
from tasks import eval_utils, finetune_utils
from tasks.race.data import RaceDataset


def train_valid_datasets_provider():
    args, tok = get_args(), get_tokenizer()
    return (RaceDataset("training", args.train_data, tok, args.seq_length),
            RaceDataset("validation", args.valid_data, tok, args.seq_length))


def model_provider(pre_process=True, post_process=True):
    print_rank_0("building multichoice model for RACE ...")
    return MultipleChoice(num_tokentypes=2, pre_process=pre_process, post_process=post_process,
                          model_type=ModelType.encoder_or_decoder)


def metrics_func_provider():
    args, tok = get_args(), get_tokenizer()

    def single_dataset_provider(datapath):
        name = datapath.split("RACE")[-1].strip("/").replace("/", "-")
        return RaceDataset(name, [datapath], tok, args.seq_length)
    return eval_utils.accuracy_func_provider(single_dataset_provider)


def main():
    finetune_utils.finetune(train_valid_datasets_provider, model_provider, ModelType.encoder_or_decoder,
                            end_of_epoch_callback_provider=metrics_func_provider)
