"""Find the first module whose output is non-finite in the smoke configuration (GPU debug helper)."""
import os, sys, torch
sys.path.insert(0, ".")
os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29544")
import finetune
from megatron_llm_b200 import get_args
from megatron_llm_b200.initialize import initialize_megatron
from megatron_llm_b200.models import ModelType
from megatron_llm_b200.training import setup_model_and_optimizer, train_step
seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
argv = (f"--model_name llama2 --num_layers 2 --hidden_size 1024 --num_attention_heads 8 --num_attention_heads_kv 8 "
        f"--ffn_hidden_size 2816 --seq_length {seq} --max_position_embeddings {seq} --micro_batch_size 1 "
        "--global_batch_size 2 --train_iters 2 --lr 1e-4 --bf16 --use_flash_attn --use_rms_norm "
        "--glu_activation swiglu --no_tie_embed_logits --position_embedding_type rotary --hidden_dropout 0 "
        "--attention_dropout 0 --tokenizer_type NullTokenizer --vocab_file 32000 --data_type synthetic "
        "--log_interval 1 --eval_iters 0 --eval_interval 1000 --num_workers 0 --lr_decay_style constant").split()
initialize_megatron(finetune.extra_args, {}, args_list=argv)
model, opt, sched = setup_model_and_optimizer(finetune.model_provider, ModelType.encoder_or_decoder)
bad = []
def mk(name):
    def hook(mod, inp, out):
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for i, o in enumerate(outs):
            if isinstance(o, torch.Tensor) and o.is_floating_point() and not torch.isfinite(o).all():
                bad.append((name, i, tuple(o.shape)))
        ins = inp if isinstance(inp, (tuple, list)) else (inp,)
        for i, o in enumerate(ins):
            if isinstance(o, torch.Tensor) and o.is_floating_point() and not torch.isfinite(o).all():
                bad.append((name + ":INPUT", i, tuple(o.shape)))
    return hook
for n, m in model[0].named_modules():
    m.register_forward_hook(mk(n))
for n, p in model[0].named_parameters():
    if not torch.isfinite(p).all():
        print("NONFINITE PARAM", n)
vocab = 32000
def it():
    g = torch.Generator().manual_seed(0)
    while True:
        yield {"text": torch.randint(0, vocab, (1, seq + 1), generator=g)}
loss, *_ = train_step(finetune.forward_step, it(), model, opt, sched)
torch.cuda.synchronize()
print("loss", loss)
print("first bad modules:", bad[:12])
for n, p in model[0].named_parameters():
    g = p.main_grad
    if not torch.isfinite(g).all():
        print("NONFINITE GRAD", n, tuple(g.shape), int((~torch.isfinite(g)).sum()))
print("grad norm", opt._grad_norm)
