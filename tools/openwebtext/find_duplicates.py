"""Near-duplicate detection with MinHash + banded LSH over character 5-gram shingles
(parity: tools/openwebtext/find_duplicates.py, which uses the external ``LSH`` package; the min-hashing here is a
vectorised numpy implementation).  Output: json lines ``{url: [{other_url: jaccard}, ...]}``."""
import argparse
import json
import pickle
import time
import zlib
from collections import defaultdict

import numpy as np

_PRIME = (1 << 61) - 1


def shingles(text, char_ngram=5):
    return set(text[i:i + char_ngram] for i in range(0, len(text) - char_ngram + 1)) or {text}


def jaccard(set_a, set_b, args=None):
    if not set_a or not set_b:
        return 0.0
    inter = len(set_a & set_b)
    mode = getattr(args, "jaccard", "union")
    denom = {"min": min(len(set_a), len(set_b)), "max": max(len(set_a), len(set_b))}.get(mode, len(set_a | set_b))
    return inter / denom


class MinHasher:
    def __init__(self, num_seeds, seed, char_ngram=5):
        rng = np.random.RandomState(seed)
        self.a = rng.randint(1, 1 << 31, size=num_seeds).astype(np.uint64)
        self.b = rng.randint(0, 1 << 31, size=num_seeds).astype(np.uint64)
        self.char_ngram = char_ngram

    def fingerprint(self, text):
        h = np.fromiter((zlib.crc32(s.encode("utf-8")) for s in shingles(text, self.char_ngram)), dtype=np.uint64)
        vals = (h[None, :] * self.a[:, None] + self.b[:, None]) % np.uint64(_PRIME)
        return vals.min(axis=1)


def compute_fingerprint(line, key, hasher):
    try:
        doc = json.loads(line)
        return doc[key], doc["text"], hasher.fingerprint(doc["text"]), True
    except Exception as e:
        print("Error:", e)
        return None, None, None, False


def candidate_buckets(fingerprints, num_bands):
    """band -> bucket hash -> urls."""
    tables = [defaultdict(list) for _ in range(num_bands)]
    for url, fp in fingerprints.items():
        for b, band in enumerate(np.array_split(fp, num_bands)):
            tables[b][band.tobytes()].append(url)
    return tables


def url_pairs_to_remove(args, bucket_urls, url_doc, rng):
    """Greedy within a bucket: pick a main url, drop everything similar to it, repeat ``heuristic_iter`` times."""
    out = []
    urls = list(bucket_urls)
    for _ in range(args.heuristic_iter):
        if len(urls) <= 1:
            break
        main = urls[rng.randint(len(urls))] if args.heuristic_iter > 1 else urls[0]
        main_sh = shingles(url_doc[main])
        similar = []
        for other in urls:
            if other == main:
                continue
            j = jaccard(main_sh, shingles(url_doc[other]), args)
            if j > 0.5:
                similar.append({other: j})
        if similar:
            out.append({main: similar})
        gone = {main} | {next(iter(d)) for d in similar}
        urls = [u for u in urls if u not in gone]
    return out


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--seed", type=int, default=1234, help="Random seed used for python, numpy")
    p.add_argument("--inputs", nargs="*", default=None, help="Pairwise list of the input files and keys, "
                   "e.g. --inputs cc.json cc_id news.json news_id")
    p.add_argument("--load_fingerprints", nargs="*", default=None, help="Load fingerprints from a list of pickle files")
    p.add_argument("--save_fingerprints", type=str, default=None, help="Save the fingerprints of the inputs")
    p.add_argument("--output", type=str, default=None, help="Output file name that consists of all ids with matching similarities")
    p.add_argument("--jaccard", type=str, default="union", choices=["union", "min", "max"], help="Jaccard similarity computation")
    p.add_argument("--heuristic_iter", type=int, default=1, help="Number of iterations to run the heuristics: use -1 for exact")
    p.add_argument("--num_bands", type=int, default=10, help="Number of bands to use in cache")
    p.add_argument("--num_seeds", type=int, default=100, help="Number of seeds to use for minhash. Note that this value should be divisible by num_bands")
    p.add_argument("--jaccard_parallel", action="store_true", help="(accepted for CLI parity; bucket scoring is cheap here)")
    args = p.parse_args()
    if args.heuristic_iter < 0:
        args.heuristic_iter = 1 << 30
    rng = np.random.RandomState(args.seed)
    hasher = MinHasher(args.num_seeds, args.seed)
    fingerprints, url_doc = {}, {}
    t0 = time.time()
    for fname in args.load_fingerprints or []:
        with open(fname, "rb") as f:
            saved = pickle.load(f)
        fingerprints.update(saved["fingerprints"])
        url_doc.update(saved["url_doc"])
    if args.inputs:
        assert len(args.inputs) % 2 == 0
        for fname, key in zip(args.inputs[::2], args.inputs[1::2]):
            with open(fname, "r", encoding="utf-8") as f:
                for n, line in enumerate(f, 1):
                    url, text, fp, ok = compute_fingerprint(line, key, hasher)
                    if ok:
                        fingerprints[url], url_doc[url] = fp, text
                    if n % 10000 == 0:
                        print(" [read]> processed {} documents in {:.2f} seconds ...".format(n, time.time() - t0), flush=True)
    if args.save_fingerprints:
        with open(args.save_fingerprints, "wb") as f:
            pickle.dump({"fingerprints": fingerprints, "url_doc": url_doc}, f)
    if args.output:
        done, n_pairs = set(), 0
        with open(args.output, "wb") as out:
            for table in candidate_buckets(fingerprints, args.num_bands):
                for bucket in table.values():
                    bucket = [u for u in dict.fromkeys(bucket) if u not in done]
                    if len(bucket) <= 1:
                        continue
                    for group in url_pairs_to_remove(args, bucket, url_doc, rng):
                        main = next(iter(group))
                        done.add(main)
                        done.update(next(iter(d)) for d in group[main])
                        out.write(json.dumps(group, ensure_ascii=False).encode("utf-8") + b"\n")
                        n_pairs += len(group[main])
        print("found {} near-duplicate pairs in {:.2f} seconds".format(n_pairs, time.time() - t0), flush=True)
    print("done :-)")
