"""Re-derives tests/golden/readme_example.json's blob bytes with the C oracle and checks them against
the committed fixture (whose query answers are the reference's own README goldens).
Run from the repo root: python tests/golden/make_readme_example.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as po  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
g = json.load(open(os.path.join(here, "readme_example.json")))
tbl, sc = po.encoding_table([s.encode() for s in g["symbols"]])
t = po.IndexType(**g["type"])
blob = po.build_blob(t, g["text"].encode(), sc, tbl, g["kmer_size"], g["sampling_ratio"])
body = blob[g["header_size"]:].tobytes().hex()
print("blob_size", blob.size, "matches" if blob.size == g["blob_size"] else "DIFFERS")
print("body_hex", "matches" if body == g["body_hex"] else "DIFFERS")
print(body)
