"""Where does the process leave the capture side stream?  (debug aid for --cuda_graph_microbatch)"""
import os, sys, json, torch
sys.path.insert(0, ".")
os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29699")
import tests.test_model_gpu as T
import finetune
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.training import setup_model_and_optimizer, train_step
import megatron_llm_b200.training as tr

argv = T.CONFIGS["llama"].replace("--train_iters 10", "--train_iters 20")
if len(sys.argv) > 1 and sys.argv[1] == "gpuinit":
    argv = argv.replace("--use_cpu_initialization", "")
argv = (argv + " --cuda_graph_microbatch").split()


def where(tag):
    print(f"STREAM {tag}: default={torch.cuda.current_stream() == torch.cuda.default_stream()} {torch.cuda.current_stream()}", flush=True)


initialize_megatron(finetune.extra_args, {}, args_list=argv)
where("after initialize_megatron")
orig_get_model = tr.get_model
def traced_get_model(*a, **k):
    where("get_model enter")
    m = orig_get_model(*a, **k)
    where("get_model exit")
    return m
tr.get_model = traced_get_model
model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
where("after setup_model_and_optimizer")
g = torch.Generator().manual_seed(0)
batches = [torch.randint(0, 1000, (2, 257), generator=g) for _ in range(2)]
def it():
    while True:
        for b in batches:
            yield {"text": b}
data = it()
for step in range(3):
    try:
        loss, skipped, gnorm, _ = train_step(finetune.forward_step, data, model, opt, sched)
        where(f"after step {step} loss {loss['lm loss'].item():.4f}")
    except Exception as e:
        where(f"step {step} raised {type(e).__name__}: {str(e)[:200]}")
        break
