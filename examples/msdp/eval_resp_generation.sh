#!/bin/bash
# Token-level F1 of the generated responses (parity: examples/msdp/eval_resp_generation.sh).
source "$(dirname "$0")/../_common.sh"; DIR=${DIR:-data/msdp}
GPUS_PER_NODE=1 launch $REPO/tasks/msdp/main.py --tensor_model_parallel_size 1 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --seq_length 2048 --max_position_embeddings 2048 --micro_batch_size 1 --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} --merge_file ${MERGE_FILE:-gpt2-merges.txt} --tokenizer_type GPT2BPETokenizer --bf16 --task MSDP-EVAL-F1 --guess_file $DIR/wow/resp_gen_seen.txt \
  --answer_file $DIR/wow/test_seen_resp_ref.txt
