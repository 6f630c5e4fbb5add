"""tools/openwebtext: cleaning, MinHash dedup, grouping, n-gram decontamination on a toy corpus."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
OWT = os.path.join(ROOT, "tools", "openwebtext")


def _run(script, *argv):
    r = subprocess.run([sys.executable, os.path.join(OWT, script), *argv], capture_output=True, text=True, cwd=OWT,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def _jsonl(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


def _read(path):
    with open(path) as f:
        return [json.loads(l) for l in f if l.strip()]


EN = ("The quick brown fox jumps over the lazy dog and then it runs to the river where all of the other animals "
      "have been waiting for it since the morning. ") * 12


def test_cleanup_and_dedup_pipeline(tmp_path):
    docs = [{"text": EN + "First unique ending about mountains.", "url": "http://a.com/1"},
            {"text": EN + "First unique ending about mountains!", "url": "http://b.com/2"},     # near duplicate
            {"text": "Der schnelle braune Fuchs springt über den faulen Hund und läuft zum Fluss. " * 30,
             "url": "http://c.de/3"},
            {"text": "too short", "url": "http://d.com/4"},
            {"text": ("Completely different content about compilers, tensor cores and memory hierarchies that "
                      "shares nothing with the other documents in this tiny corpus at all. ") * 15,
             "url": "http://e.com/5"}]
    _jsonl(tmp_path / "raw.json", docs)
    _run("cleanup_dataset.py", str(tmp_path / "raw.json"), str(tmp_path / "clean.json"))
    clean = _read(tmp_path / "clean.json")
    assert [d["url"] for d in clean] == ["http://a.com/1", "http://b.com/2", "http://e.com/5"]
    _run("find_duplicates.py", "--inputs", str(tmp_path / "clean.json"), "url", "--output", str(tmp_path / "pairs.json"))
    pairs = _read(tmp_path / "pairs.json")
    assert len(pairs) == 1 and {next(iter(pairs[0]))} | {next(iter(d)) for d in next(iter(pairs[0].values()))} == \
        {"http://a.com/1", "http://b.com/2"}
    _run("group_duplicate_url.py", str(tmp_path / "pairs.json"), str(tmp_path / "groups.json"), "0.7")
    _run("remove_group_duplicates.py", str(tmp_path / "groups.json"), str(tmp_path / "clean.json"),
         str(tmp_path / "dedup.json"))
    assert len(_read(tmp_path / "dedup.json")) == 2
    _run("add_id.py", "--input_file", str(tmp_path / "dedup.json"), "--output_file", str(tmp_path / "ids.json"),
         "--id_prefix", "owt")
    assert _read(tmp_path / "ids.json")[1]["adlr_id"] == "owt-0000000002"


def test_filter_ngrams(tmp_path):
    leak = "alpha beta gamma delta epsilon zeta eta theta iota kappa lambda mu nu xi"
    _jsonl(tmp_path / "task.jsonl", [{"text": leak}])
    filler = "This sentence is ordinary filler text that talks about nothing in particular. " * 8
    _jsonl(tmp_path / "corpus.json", [{"text": filler + leak + ". " + filler, "id": 1}, {"text": filler * 2, "id": 2}])
    _run("filter_ngrams.py", "--tasks", str(tmp_path / "task.jsonl"), "--dedup_dataset", str(tmp_path / "corpus.json"),
         "text", "--output", str(tmp_path / "out.json"))
    out = _read(tmp_path / "out.json")
    assert all("gamma delta" not in d["text"] for d in out)
    assert sorted(d["id"] for d in out) == [1, 1, 2] and sum("split_id" in d for d in out) == 2


def test_url_blacklist(tmp_path):
    (tmp_path / "urls").mkdir()
    (tmp_path / "urls" / "u.txt").write_text("\n".join([
        "http://example.com/article/one", "http://example.com/article/one", "https://www.youtube.com/watch?v=1",
        "http://example.org/file.pdf", "notaurl", "http://news.site.co.uk/story"]) + "\n")
    _run("blacklist_urls.py", str(tmp_path / "urls"), str(tmp_path / "clean.txt"))
    kept = set((tmp_path / "clean.txt").read_text().split())
    assert kept == {"http://example.com/article/one", "http://news.site.co.uk/story"}
