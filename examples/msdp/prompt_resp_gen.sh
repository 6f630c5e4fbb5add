#!/bin/bash
# Stage 2: response generation conditioned on the generated knowledge (parity: examples/msdp/prompt_resp_gen.sh).
source "$(dirname "$0")/../_common.sh"; DIR=${DIR:-data/msdp}
launch $REPO/tasks/msdp/main.py --tensor_model_parallel_size 1 --num_layers 24 --hidden_size 1024 --num_attention_heads 16 --seq_length 2048 --max_position_embeddings 2048 --micro_batch_size 1 --vocab_file ${VOCAB_FILE:-gpt2-vocab.json} --merge_file ${MERGE_FILE:-gpt2-merges.txt} --tokenizer_type GPT2BPETokenizer --bf16 --load ${CHECKPOINT_PATH:-checkpoints/gpt2_345m} --task MSDP-PROMPT --prompt_type response \
  --sample_input_file $DIR/wow/resp_gen_input_seen.txt --sample_output_file $DIR/wow/resp_gen_seen.txt \
  --prompt_file $DIR/wow/resp_prompts.txt --out_seq_length 100 --num_prompt_examples 20
