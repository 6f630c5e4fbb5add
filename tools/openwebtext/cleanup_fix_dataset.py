"""Dataset-specific filters and fixes (parity: tools/openwebtext/cleanup_fix_dataset.py).

--tasks: remove_512 | remove_256_javascript | remove_512_non_english | ftfy_fix_text | general_cleaning.
Writes ``<output_path>/<name>_cleaned.json`` and ``_filtered.json`` (the dropped documents) per input file."""
import argparse
import json
import os
import re
import time

from textutils import detect_language, fix_text, write_jsonl_row

TASKS = ["remove_512", "remove_256_javascript", "remove_512_non_english", "ftfy_fix_text", "general_cleaning"]


def process_doc(json_line, args):
    """-> (which task fired, text, document, drop?)."""
    doc = json.loads(json_line)
    text = doc["text"]
    fired = dict.fromkeys(TASKS, False)
    try:
        filters = [("remove_512", lambda t: len(t) < 512),
                   ("remove_256_javascript", lambda t: len(t) < 256 and "javascript" in t.lower()),
                   ("remove_512_non_english", lambda t: len(t) < 512 and detect_language(t) != "en")]
        for name, cond in filters:
            if name in args.tasks and cond(text):
                fired[name] = True
                return fired, text, doc, True
        if "ftfy_fix_text" in args.tasks:
            fired["ftfy_fix_text"] = True
            return fired, fix_text(text), doc, False
        if "general_cleaning" in args.tasks:
            fired["general_cleaning"] = True
            return fired, re.sub(r"  +|\b\n+ |\b\n+", " ", text), doc, False
    except Exception as e:
        print("Error: *************************\n{}\ntext: {}".format(e, text), flush=True)
        return fired, text, doc, True
    return fired, text, doc, False


def process_set(args, input_file, out_cleaned, out_filtered):
    counts = dict.fromkeys(TASKS, 0)
    t0 = time.time()
    with open(input_file, "r", encoding="utf-8") as fin, open(out_cleaned, "wb") as fc, open(out_filtered, "wb") as ff:
        for n, line in enumerate(fin, 1):
            fired, text, doc, drop = process_doc(line, args)
            for k, v in fired.items():
                counts[k] += v
            if drop:
                write_jsonl_row(ff, doc)
            else:
                doc["text"] = text
                write_jsonl_row(fc, doc)
            if n % args.log_interval == 0:
                print("    processed {:9d} documents in {:.2f} seconds ...".format(n, time.time() - t0), flush=True)
    print("{}: {}".format(input_file, counts), flush=True)


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--input_files", nargs="*", required=True, default=None, help="Input json files that needs to be cleaned")
    p.add_argument("--tasks", nargs="*", required=True, default=None, help="Tasks to perform: " + ", ".join(TASKS))
    p.add_argument("--output_path", type=str, default=None, help="Directory where the output should go")
    p.add_argument("--log_interval", type=int, default=100, help="Log interval")
    args = p.parse_args()
    for f in args.input_files:
        stem = os.path.splitext(os.path.basename(f))[0]
        process_set(args, f, os.path.join(args.output_path, stem + "_cleaned.json"),
                    os.path.join(args.output_path, stem + "_filtered.json"))
    print("done :-)", flush=True)
