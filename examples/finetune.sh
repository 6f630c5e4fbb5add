#!/bin/bash
# Pre-train / fine-tune / instruction-tune llama, llama2, codellama, falcon, mistral or gpt (parity: examples/finetune.sh).
#   examples/finetune.sh llama2 --size 7 --tp 8 --checkpoint /ckpts/llama2-7b-tp8-pp1 --data /data/corpus_text_document
source "$(dirname "$0")/_common.sh"
SIZE=7; TP=8; PP=1; MICRO_BATCH=1; GLOBAL_BATCH=12; WANDB=0; INSTRUCT=0; CHECKPOINT_PATH=none; DATA_PATH=none
TRAINED_PATH=none; VAL_PATH=none; ITERS=1000; SEQ_LEN=none; USR_LR=none; USR_MIN_LR=none; LOSS_MASK=0.0
TOKENIZER_MODEL=${TOKENIZER_MODEL:-/models/llama2/tokenizer.model}
help () { echo "Usage: $0 <gpt/llama/llama2/codellama/falcon/mistral> [--size N] [--tp N] [--pp N] [--gpus N] [--micro-batch N] [--global-batch N] [--nodes N] [--rank N] [--addr A] [--wandb] [--instruct] [--checkpoint P] [--data P] [--iters N] [--seq-len N] [--val-path P] [--out P] [--lr LR MINLR] [--loss-mask X]"; }
[[ $# = 0 ]] && { help; exit 1; }
[[ $1 = -h || $1 = --help ]] && { help; exit 0; }
MODEL=$1; shift
while [[ $# -gt 0 ]]; do case $1 in
  --tp) TP=$2; shift 2;; --pp) PP=$2; shift 2;; --size) SIZE=$2; shift 2;; --gpus) GPUS_PER_NODE=$2; shift 2;;
  --micro-batch) MICRO_BATCH=$2; shift 2;; --global-batch) GLOBAL_BATCH=$2; shift 2;; --rank) NODE_RANK=$2; shift 2;;
  --nodes) NNODES=$2; shift 2;; --addr) MASTER_ADDR=$2; shift 2;; --wandb) WANDB=1; shift;; --instruct) INSTRUCT=1; shift;;
  --checkpoint) CHECKPOINT_PATH=$2; shift 2;; --data) DATA_PATH=$2; shift 2;; --iters) ITERS=$2; shift 2;;
  --seq-len) SEQ_LEN=$2; shift 2;; --out) TRAINED_PATH=$2; shift 2;; --val-path) VAL_PATH=$2; shift 2;;
  --lr) USR_LR=$2; USR_MIN_LR=$3; shift 3;; --loss-mask) LOSS_MASK=$2; shift 2;; *) echo "unknown argument $1"; help; exit 1;;
esac; done
source "$(dirname "$0")/_common.sh"   # re-evaluate DISTRIBUTED_ARGS with the parsed values
[[ $CHECKPOINT_PATH = none ]] && CHECKPOINT_PATH=/checkpoints/${MODEL}-${SIZE}b-tp$TP-pp$PP
if [[ $INSTRUCT = 1 ]]; then LR=2e-5; MIN_LR=2e-6; SUFFIX=instructed; else LR=3e-4; MIN_LR=3e-4; SUFFIX=pretrained; fi
[[ $TRAINED_PATH = none ]] && TRAINED_PATH=$CHECKPOINT_PATH-$SUFFIX
[[ $USR_LR != none ]] && { LR=$USR_LR; MIN_LR=$USR_MIN_LR; }
case $MODEL in
  falcon) TOKENIZER=FalconTokenizer; EXTRA_ARGS="--parallel_attn"; [[ $SEQ_LEN = none ]] && SEQ_LEN=2048
          (( SIZE > 7 )) && EXTRA_ARGS="$EXTRA_ARGS --parallel_layernorm";;
  llama|llama2|codellama|mistral)
          TOKENIZER=SentencePieceTokenizer
          EXTRA_ARGS="--vocab_file=$TOKENIZER_MODEL --use_rms_norm --glu_activation swiglu --no_tie_embed_logits"
          [[ $INSTRUCT = 1 ]] && EXTRA_ARGS="$EXTRA_ARGS --vocab_extra_ids_list <|im_start|>,<|im_end|>"
          case $MODEL in
            llama) [[ $SEQ_LEN = none ]] && SEQ_LEN=2048; EXTRA_ARGS="$EXTRA_ARGS --layernorm_epsilon 1e-6";;
            llama2) [[ $SEQ_LEN = none ]] && SEQ_LEN=4096; EXTRA_ARGS="$EXTRA_ARGS --layernorm_epsilon 1e-5"; (( SIZE > 13 )) && LR=1.5e-4;;
            codellama) [[ $SEQ_LEN = none ]] && SEQ_LEN=16384; EXTRA_ARGS="$EXTRA_ARGS --rope_theta 1e6";;
            mistral) [[ $SEQ_LEN = none ]] && SEQ_LEN=8192; EXTRA_ARGS="$EXTRA_ARGS --sliding_window_size 4096";;
          esac;;
  gpt)    TOKENIZER=GPT2BPETokenizer; [[ $SEQ_LEN = none ]] && SEQ_LEN=2048
          EXTRA_ARGS="--num_layers 4 --hidden_size 512 --num_attention_heads 8 --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} --merge_file ${MERGE_FILE:-gpt2-merges.txt}";;
  *) echo "Model should be either gpt, llama, llama2, codellama, falcon or mistral, not $MODEL"; help; exit 1;;
esac
COMMON="$COMMON_ARGS --tensor_model_parallel_size $TP --pipeline_model_parallel_size $PP --load $CHECKPOINT_PATH --save $TRAINED_PATH
  --tensorboard_dir $TRAINED_PATH/logging --data_path $DATA_PATH --model_name $MODEL --tokenizer_type $TOKENIZER
  --global_batch_size $GLOBAL_BATCH --micro_batch_size $MICRO_BATCH --sequence_parallel --recompute_granularity selective
  --use_checkpoint_args --use_distributed_optimizer --seq_length $SEQ_LEN --train_iters $ITERS --lr $LR --min_lr $MIN_LR
  --log_interval 10 --save_interval 100 --eval_interval 100 --eval_iters 10 --hidden_dropout 0.0 --position_embedding_type rotary
  --lr_decay_style cosine --lr_warmup_iters 50 --weight_decay 0.1 --clip_grad 1.0 --no_new_tokens --adam_beta1 0.9 --adam_beta2 0.95
  --log_timers_to_tensorboard --tensorboard_log_interval 10 --scalar_loss_mask $LOSS_MASK --split 989,10,1"
[[ $INSTRUCT = 1 ]] && COMMON="$COMMON --variable_seq_lengths --data_type instruction --metrics all"
[[ $INSTRUCT = 1 ]] || COMMON="$COMMON --metrics perplexity accuracy count_loss_mask"
[[ $VAL_PATH != none ]] && COMMON="$COMMON --train_data_path $DATA_PATH --valid_data_path $VAL_PATH"
[[ $WANDB = 1 ]] && COMMON="$COMMON --wandb_logger"
[[ "$(cat $CHECKPOINT_PATH/latest_checkpointed_iteration.txt 2>/dev/null)" = release ]] && COMMON="$COMMON --finetune"
launch $REPO/finetune.py $COMMON $EXTRA_ARGS
