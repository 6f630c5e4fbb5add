"""Launch the REST text-generation server for any supported model family.

Parity: tools/run_text_generation_server.py.  The reference hard-codes GPTModel; this launcher takes ``--model_name``
(gpt | llama | llama2 | codellama | falcon | mistral) like finetune.py so a Llama/Falcon checkpoint can be served.
Rank (pp=0, tp=0) runs the HTTP server; all other ranks loop on the broadcast op code."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from megatron_llm_b200 import get_args, print_rank_0  # noqa: E402
from megatron_llm_b200.checkpointing import load_checkpoint  # noqa: E402
from megatron_llm_b200.initialize import initialize_megatron  # noqa: E402
from megatron_llm_b200.models.enums import ModelType  # noqa: E402
from megatron_llm_b200.parallel import state as mpu  # noqa: E402
from megatron_llm_b200.text_generation import beam_search_and_post_process, generate_and_post_process  # noqa: E402
from megatron_llm_b200.text_generation_server import BEAM_NUM, GENERATE_NUM, MegatronServer  # noqa: E402
from megatron_llm_b200.training import get_model  # noqa: E402
from megatron_llm_b200.utils.device import current_device  # noqa: E402


def model_provider(pre_process=True, post_process=True):
    import finetune
    print_rank_0("building model for generation ...")
    model = finetune.model_provider(pre_process, post_process)
    from megatron_llm_b200.utils import unwrap_model
    unwrap_model(model).parallel_output = False
    return model


def add_text_generate_args(parser):
    group = parser.add_argument_group(title="text generation")
    group.add_argument("--temperature", type=float, default=1.0, help="Sampling temperature.")
    group.add_argument("--top_p", type=float, default=0.0, help="Top p sampling.")
    group.add_argument("--top_k", type=int, default=0, help="Top k sampling.")
    group.add_argument("--out_seq_length", type=int, default=1024, help="Size of the output generated text.")
    group.add_argument("--model_name", default="gpt",
                       choices={"gpt", "llama", "falcon", "llama2", "codellama", "mistral"})
    group.add_argument("--model_type", default="encoder_or_decoder")
    group.add_argument("--port", type=int, default=5000)
    return parser


def main(args_list=None):
    initialize_megatron(extra_args_provider=add_text_generate_args,
                        args_defaults={"tokenizer_type": "GPT2BPETokenizer", "no_load_rng": True,
                                       "no_load_optim": True}, args_list=args_list)
    args = get_args()
    if args.num_layers_per_virtual_pipeline_stage is not None:
        print("Interleaved pipeline schedule is not yet supported for text generation.")
        sys.exit()
    model = get_model(model_provider, ModelType.encoder_or_decoder, wrap_with_ddp=False, args=args)
    if args.load is not None:
        load_checkpoint(model, None, None)
    assert len(model) == 1, "Above condition should have caught this"
    model = model[0]
    if mpu.is_pipeline_first_stage() and mpu.get_tensor_model_parallel_rank() == 0:
        MegatronServer(model).run("0.0.0.0", port=args.port)
        return
    while True:
        choice = torch.zeros(1, dtype=torch.long, device=current_device())
        dist.broadcast(choice, 0)
        try:
            if choice.item() == GENERATE_NUM:
                generate_and_post_process(model)
            elif choice.item() == BEAM_NUM:
                beam_search_and_post_process(model)
        except ValueError:
            pass


if __name__ == "__main__":
    main()
