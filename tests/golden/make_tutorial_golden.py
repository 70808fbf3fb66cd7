#!/usr/bin/env python3
"""Extract the documentation goldens of the reference into tests/golden/tutorial.json.

Source: /root/reference/tutorial.ipynb (outputs of cells run by the reference's own Python binding
on examples/eng.aspell.lexicon + examples/simple.alphabet.tsv with default SearchParameters).
Cells used: find_variants("separate"), find_variants("seperate"),
find_all_matches("We would like seperate beds"), find_all_matches("We would like sep arate beds")[3].
Run here (the reference tree does not exist on the GPU box); the JSON is committed.
"""
import ast
import json
import os
import sys

NB = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tutorial.ipynb"
nb = json.load(open(NB, encoding="utf-8"))
out = {"find_variants": {}, "find_all_matches": {}}
for cell in nb["cells"]:
    if cell["cell_type"] != "code":
        continue
    src = "".join(cell["source"])
    text = "".join("".join(o.get("text", [])) for o in cell.get("outputs", []) if o.get("name") == "stdout")
    if not text:
        continue
    rows = [ast.literal_eval(line) for line in text.splitlines() if line.startswith("{")]
    if 'model.find_variants("separate"' in src:
        out["find_variants"]["separate"] = rows
    elif 'model.find_variants("seperate"' in src:
        out["find_variants"]["seperate"] = rows
    elif 'model.find_all_matches("We would like seperate beds"' in src:
        out["find_all_matches"]["We would like seperate beds"] = rows
    elif 'model.find_all_matches("We would like sep arate beds"' in src:
        out["find_all_matches"]["We would like sep arate beds"] = {"index": 3, "match": rows[0]}
for group in out.values():
    for v in group.values():
        assert v, "golden cell not found"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tutorial.json")
json.dump(out, open(dst, "w", encoding="utf-8"), indent=1, ensure_ascii=False)
print("wrote", dst, {k: list(v) for k, v in out.items()})
