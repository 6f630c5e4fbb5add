#!/bin/bash
# GPT-3 175B: TP=8 (one NVSwitch domain), PP=16, interleaved schedule, batch ramp-up (parity: examples/pretrain_gpt3_175B.sh).
#SBATCH --nodes=128 --exclusive --ntasks-per-node=8 --job-name=gpt3_175b
source "$(dirname "$0")/_common.sh"
DIR=${DIR:-$(pwd)}; DATASET="${DATASET:-0.5 ds1_text_document 0.5 ds2_text_document}"
srun python -u $REPO/finetune.py --model_name gpt --tensor_model_parallel_size 8 --pipeline_model_parallel_size 16 \
  --num_layers_per_virtual_pipeline_stage 2 --sequence_parallel --num_layers 96 --hidden_size 12288 --num_attention_heads 96 \
  --seq_length 2048 --max_position_embeddings 2048 --micro_batch_size 1 --global_batch_size 1536 --rampup_batch_size 16 16 5859375 \
  --train_samples 146484375 --lr_decay_samples 126953125 --lr_warmup_samples 183105 --lr 6.0e-5 --min_lr 6.0e-6 --lr_decay_style cosine \
  --log_interval 10 --eval_iters 40 --eval_interval 1000 --data_path $DATASET --vocab_file gpt2-vocab.json --merge_file gpt2-merges.txt \
  --tokenizer_type GPT2BPETokenizer --save_interval 1000 --save $DIR/checkpoints --load $DIR/checkpoints --split 98,2,0 --clip_grad 1.0 \
  --weight_decay 0.1 --adam_beta1 0.9 --adam_beta2 0.95 --init_method_std 0.006 --tensorboard_dir $DIR/logs --bf16 \
  --recompute_granularity selective --use_distributed_optimizer --DDP_impl local
