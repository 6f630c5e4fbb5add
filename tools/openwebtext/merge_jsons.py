"""Concatenate every ``*.json`` (json-lines) file of a directory (parity: tools/openwebtext/merge_jsons.py)."""
import argparse
import glob
import json

if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--json_path", type=str, default=".", help="path where all the json files are located")
    p.add_argument("--output_file", type=str, default="merged_output.json", help="filename of the merged json")
    args = p.parse_args()
    with open(args.output_file, "w") as out:
        for n, fname in enumerate(sorted(glob.glob(args.json_path + "/*.json")), 1):
            if n % 1024 == 0:
                print("Merging at ", n, flush=True)
            with open(fname, "r") as f:
                for row in f:
                    json.loads(row)         # validate
                    out.write(row if row.endswith("\n") else row + "\n")
    print("Merged file", args.output_file, flush=True)
