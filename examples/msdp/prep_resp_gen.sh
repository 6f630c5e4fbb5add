#!/bin/bash
# Put the generated knowledge into the response-generation inputs (parity: examples/msdp/prep_resp_gen.sh).
REPO=$(cd "$(dirname "$0")/../.." && pwd); DIR=${DIR:-data/msdp}; export PYTHONPATH=$REPO
python $REPO/tasks/msdp/preprocessing.py --func prepare_input --test_file $DIR/wow/test_seen_processed.txt \
  --knwl_gen_file $DIR/wow/knwl_gen_seen.txt --processed_file $DIR/wow/resp_gen_input_seen.txt
