"""First cleaning pass: fix broken unicode, keep English documents with at least 128 tokens
(parity: tools/openwebtext/cleanup_dataset.py).   usage: cleanup_dataset.py <input jsonl> <output jsonl>"""
import sys
import time

from textutils import detect_language, fix_text, read_jsonl, write_jsonl_row

MIN_DOCUMENT_LENGHT = 128


def count_tokens(text):
    """GPT-2 BPE would give ~1.3 tokens / word; whitespace words are a conservative lower bound."""
    return len(text.split())


def filter_corpus(filename, out_filename, print_interval=10000):
    print(" > filtering {}".format(filename))
    stats = dict(docs=0, written=0, fixed=0, non_english=0, small=0)
    t0 = time.time()
    with open(out_filename, "wb") as out:
        for row in read_jsonl(filename):
            stats["docs"] += 1
            try:
                text = fix_text(row["text"])
                stats["fixed"] += text != row["text"]
                row["text"] = text
                if detect_language(text) != "en":
                    stats["non_english"] += 1
                    continue
                if len(text) < 8 * MIN_DOCUMENT_LENGHT and count_tokens(text) < MIN_DOCUMENT_LENGHT:
                    stats["small"] += 1
                    continue
                write_jsonl_row(out, row)
                stats["written"] += 1
            except Exception as e:
                print("    skipping ", row, e)
            if stats["docs"] % print_interval == 0:
                print("[PROGRESS] {:.1f}s {}".format(time.time() - t0, stats), flush=True)
    print("[FINAL] {:.1f}s {}".format(time.time() - t0, stats), flush=True)
    return stats


if __name__ == "__main__":
    print("building gpt2 dataset ...")
    filter_corpus(sys.argv[1], sys.argv[2])
