"""Merge the pairwise duplicate lists of find_duplicates.py into groups of mutually similar urls (union-find)
(parity: tools/openwebtext/group_duplicate_url.py).   usage: group_duplicate_url.py <pairs> <groups> [threshold=0.7]"""
import json
import sys
import time

if __name__ == "__main__":
    print("grouping duplicate urls ...")
    inp, out = sys.argv[1], sys.argv[2]
    threshold = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    parent = {}

    def find(u):
        parent.setdefault(u, u)
        while parent[u] != u:
            parent[u] = parent[parent[u]]
            u = parent[u]
        return u

    t0 = time.time()
    with open(inp, "r") as f:
        for n, line in enumerate(f, 1):
            for main, others in json.loads(line).items():
                root = find(main)
                for entry in others:
                    for url, sim in entry.items():
                        if sim >= threshold:
                            parent[find(url)] = root
            if n % 100000 == 0:
                print(" > processed {} lines in {} seconds ...".format(n, time.time() - t0))
    groups = {}
    for u in list(parent):
        groups.setdefault(find(u), []).append(u)
    groups = [g for g in groups.values() if len(g) > 1]
    remove = sum(len(g) - 1 for g in groups)
    print("out of {} urls, only {} are unique and {} should be removed".format(remove + len(groups), len(groups), remove))
    with open(out, "wb") as f:
        for i, g in enumerate(groups):
            f.write(json.dumps({str(i): g}, ensure_ascii=False).encode("utf-8") + b"\n")
