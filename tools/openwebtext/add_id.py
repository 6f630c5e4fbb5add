"""Add a unique ``adlr_id`` (``<prefix>-0000000001`` ...) to every json line (parity: tools/openwebtext/add_id.py)."""
import argparse
import time

from textutils import read_jsonl, write_jsonl_row

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--input_file", type=str, default=None, help="Input json file where id needs to be added")
    p.add_argument("--output_file", type=str, default=None, help="Output file name with id")
    p.add_argument("--id_prefix", type=str, default=None, help="Id prefix")
    p.add_argument("--log_interval", type=int, default=100, help="Log interval")
    args = p.parse_args()
    t0 = time.time()
    with open(args.output_file, "wb") as out:
        for n, row in enumerate(read_jsonl(args.input_file), 1):
            row["adlr_id"] = "{}-{:010d}".format(args.id_prefix, n)
            write_jsonl_row(out, row)
            if n % args.log_interval == 0:
                print("    processed {:9d} documents in {:.2f} seconds ...".format(n, time.time() - t0), flush=True)
    print("done :-)", flush=True)
