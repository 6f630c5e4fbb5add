"""Per-kernel time breakdown of one training step (torch.profiler, CUDA activities).  Analysis only -- numbers
taken under a profiler are never reported as benchmark values."""
import os, sys, json, torch
sys.path.insert(0, ".")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
import finetune, bench
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.training import setup_model_and_optimizer, train_step

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/profile_step.txt"
class A: pass
a = A(); a.model = "llama2-7b"; a.layers = layers; a.seq = 4096; a.micro_batch = 1; a.global_batch = 2
argv, vocab = bench.megatron_argv(a, 1)
argv += ["--tokenizer_type", "NullTokenizer", "--vocab_file", str(vocab), "--data_type", "synthetic"]
initialize_megatron(finetune.extra_args, {}, args_list=argv)
model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
def it():
    g = torch.Generator().manual_seed(0)
    while True:
        yield {"text": torch.randint(0, vocab, (1, 4097), generator=g).cuda()}
data = it()
for _ in range(2):
    train_step(finetune.forward_step, data, model, opt, sched)
torch.cuda.synchronize()
_s, _e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_s.record(); train_step(finetune.forward_step, data, model, opt, sched); _e.record(); torch.cuda.synchronize()
step_ms = _s.elapsed_time(_e)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_step(finetune.forward_step, data, model, opt, sched)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    dt = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if dt > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
        rows.append((dt, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
with open(out, "w") as f:
    f.write(f"layers={layers} global_batch=2 total_device_us={tot:.0f} unprofiled_step_ms={step_ms:.1f} "
            f"(device busy {tot / 10 / step_ms:.1f}% of the step)\n")
    for dt, n, k in rows[:45]:
        f.write(f"{dt/tot*100:6.2f}%  {dt:10.0f}us  n={n:5d}  {k[:110]}\n")
print(open(out).read())
