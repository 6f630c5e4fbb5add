"""GLUE (MNLI / QQP) fine-tuning and evaluation (parity: tasks/glue/finetune.py)."""
from megatron_llm_b200 import get_args, get_tokenizer, print_rank_0
from megatron_llm_b200.models.classification import Classification
from megatron_llm_b200.models.enums import ModelType
from tasks import finetune_utils
from tasks.eval_utils import accuracy_func_provider


def _glue_classification(num_classes, Dataset, name_from_datapath_func):
    def train_valid_datasets_provider():
        args, tok = get_args(), get_tokenizer()
        return (Dataset("training", args.train_data, tok, args.seq_length),
                Dataset("validation", args.valid_data, tok, args.seq_length))

    def model_provider(pre_process=True, post_process=True):
        print_rank_0("building classification model for {} ...".format(get_args().task))
        return Classification(num_classes=num_classes, num_tokentypes=2, pre_process=pre_process,
                              post_process=post_process, model_type=ModelType.encoder_or_decoder)

    def metrics_func_provider():
        def single_dataset_provider(datapath):
            args, tok = get_args(), get_tokenizer()
            return Dataset(name_from_datapath_func(datapath), [datapath], tok, args.seq_length)
        return accuracy_func_provider(single_dataset_provider)

    finetune_utils.finetune(train_valid_datasets_provider, model_provider, ModelType.encoder_or_decoder,
                            end_of_epoch_callback_provider=metrics_func_provider)


def main():
    args = get_args()
    if args.task == "MNLI":
        from tasks.glue.mnli import MNLIDataset as Dataset
        num_classes = 3
    elif args.task == "QQP":
        from tasks.glue.qqp import QQPDataset as Dataset
        num_classes = 2
    else:
        raise NotImplementedError("GLUE task {} is not implemented.".format(args.task))

    def name_from_datapath(datapath):
        return datapath.split(args.task)[-1].strip(".tsv").strip("/").replace("_", "-")

    _glue_classification(num_classes, Dataset, name_from_datapath)
