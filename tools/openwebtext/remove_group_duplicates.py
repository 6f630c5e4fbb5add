"""Drop all but the first url of every duplicate group from a dataset
(parity: tools/openwebtext/remove_group_duplicates.py).   usage: remove_group_duplicates.py <groups> <data> <output>"""
import json
import sys
import time

from textutils import write_jsonl_row

if __name__ == "__main__":
    url_filename, data_filename, output_filename = sys.argv[1:4]
    urls = set()
    with open(url_filename, "r") as f:
        for line in f:
            for group in json.loads(line).values():
                urls.update(group[1:])
    print("will be removing {} urls".format(len(urls)), flush=True)
    written = removed = removed_chars = 0
    t0 = time.time()
    with open(output_filename, "wb") as fout, open(data_filename, "r") as fin:
        for line in fin:
            try:
                doc = json.loads(line)
                if doc["url"] in urls:
                    removed += 1
                    removed_chars += len(doc["text"])
                    continue
                write_jsonl_row(fout, doc)
                written += 1
            except Exception as e:
                print("[SKIPPING]", line, e)
    print(" [PROCESSED] time (s): {:.2f} | written: {} | removed: {} (char: {})".format(time.time() - t0, written,
                                                                                         removed, removed_chars))
    print("done :-)")
