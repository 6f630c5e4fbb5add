#!/bin/bash
# Wizard-of-Wikipedia / Wizard-of-Internet preprocessing + prompt selection (parity: examples/msdp/data_processing.sh).
REPO=$(cd "$(dirname "$0")/../.." && pwd); DIR=${DIR:-data/msdp}; export PYTHONPATH=$REPO
P="python $REPO/tasks/msdp/preprocessing.py"
$P --func process_wow_dataset --raw_file $DIR/wow/train.json --processed_file $DIR/wow/train_processed.txt
for split in seen unseen; do
  $P --func process_wow_dataset --raw_file $DIR/wow/test_random_split.json --processed_file $DIR/wow/test_${split}_processed.txt \
     --knwl_ref_file $DIR/wow/test_${split}_knwl_ref.txt --resp_ref_file $DIR/wow/test_${split}_resp_ref.txt
done
$P --func process_woi_dataset --raw_file $DIR/woi/test.jsonl --processed_file $DIR/woi/test_processed.txt \
   --knwl_ref_file $DIR/woi/test_knwl_ref.txt --resp_ref_file $DIR/woi/test_resp_ref.txt
$P --func get_knwl_gen_prompts --test_file $DIR/wow/test_seen_processed.txt --train_file $DIR/wow/train_processed.txt \
   --model_file $DIR/dpr_question_encoder.pt --processed_file $DIR/wow/knwl_prompts_seen.jsonl --data_type wow_seen
$P --func get_resp_gen_prompts --train_file $DIR/wow/train_processed.txt --processed_file $DIR/wow/resp_prompts.txt
