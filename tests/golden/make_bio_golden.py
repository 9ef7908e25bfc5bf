"""Generates tests/golden/bio_ref.json by running the REFERENCE's own label code (TF-free):

    python tests/golden/make_bio_golden.py            # needs /root/reference (this container only)

`polus/ner/bio.py` and `polus/ner/elements.py` import nothing but each other, but `import polus` pulls TensorFlow
(polus/__init__.py:93), so the two files are loaded with stub `polus` / `polus.ner` packages in sys.modules (SURVEY.md
§8c).  Cases: seeded random token spans + aligned / misaligned / overlapping / nested / duplicate entities of two
types, plus hand-written edge cases (empty inputs, empty spans, whole-text entity, equal-length ties).  The fixture
pins `get_bio` (label indices bit-exact, §8a a23) and the reference's TAG2INT to the reference itself."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("POLUS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_bio():
    pkg, ner = types.ModuleType("polus"), types.ModuleType("polus.ner")
    pkg.__path__, ner.__path__ = [], []
    sys.modules.setdefault("polus", pkg)
    sys.modules.setdefault("polus.ner", ner)
    mods = {}
    for name in ("elements", "bio"):
        spec = importlib.util.spec_from_file_location(f"polus.ner.{name}", os.path.join(REF, "polus", "ner", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"polus.ner.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["elements"], mods["bio"]


def random_case(rng):
    n = int(rng.integers(1, 18))
    cuts = np.sort(rng.choice(np.arange(0, 120), size=2 * n, replace=False))
    spans = [(int(cuts[2 * i]), int(cuts[2 * i + 1])) for i in range(n)]
    ents = []
    for _ in range(int(rng.integers(0, 8))):
        a, b = sorted(int(v) for v in rng.integers(0, n, 2))
        s, e = spans[a][0], spans[b][1]
        u = rng.uniform()
        if u < 0.15:
            s += 1                      # left boundary inside a token: discarded
        elif u < 0.30:
            e = max(s, e - 1)           # right boundary inside a token: discarded
        elif u < 0.35 and ents:
            s, e = ents[-1][0], ents[-1][1]   # duplicate span (second one finds its tokens taken)
        if e < s:
            s, e = e, s
        ents.append((int(s), int(e), "Chemical" if rng.uniform() < 0.8 else "Gene"))
    return spans, ents


EDGE_CASES = [
    ([], []),
    ([(0, 4)], []),
    ([(0, 4)], [(0, 4, "Chemical")]),
    ([(0, 4), (5, 9), (10, 12)], [(0, 12, "Chemical")]),                               # whole text
    ([(0, 4), (5, 9), (10, 12)], [(5, 12, "Chemical"), (5, 9, "Chemical")]),            # nested: longer wins
    ([(0, 4), (5, 9), (10, 12)], [(0, 9, "Chemical"), (5, 12, "Gene")]),                # equal length overlap: first wins (stable sort)
    ([(0, 4), (5, 9), (10, 12)], [(5, 12, "Gene"), (0, 9, "Chemical")]),                # same, other order
    ([(0, 4), (4, 4), (5, 9)], [(0, 4, "Chemical"), (4, 4, "Chemical")]),               # empty token span / empty entity
    ([(3, 6), (7, 9)], [(0, 6, "Chemical")]),                                           # starts before the first token
    ([(3, 6), (7, 9)], [(3, 20, "Chemical")]),                                          # ends after the last token
    ([(0, 2), (2, 5), (5, 6)], [(0, 5, "Chemical"), (5, 6, "Chemical"), (2, 6, "Gene")]),  # adjacent tokens, no gap
]


def main():
    elements, bio = load_reference_bio()
    rng = np.random.default_rng(20261018)
    cases = list(EDGE_CASES) + [random_case(rng) for _ in range(400)]
    out = []
    for spans, ents in cases:
        objs = [elements.Entity("x" * (e - s), (s, e), t) for s, e, t in ents]
        tags = bio.get_bio([tuple(sp) for sp in spans], objs)
        out.append({"spans": [list(sp) for sp in spans], "entities": [list(e) for e in ents], "tags": tags})
    # the literal label dictionary (polus/ner/utils.py:9-15) is read from the source text: importing that module needs TF
    src = open(os.path.join(REF, "polus", "ner", "utils.py")).read()
    start = src.index("TAG2INT")
    block = src[src.index("{", start):src.index("}", start) + 1]
    tag2int = eval(block, {})  # a dict literal of str -> int
    doc = {"generator": "tests/golden/make_bio_golden.py", "reference": "polus/ner/bio.py:92-114 get_bio, polus/ner/utils.py:9-15 TAG2INT",
           "tag2int": tag2int, "cases": out}
    with open(os.path.join(HERE, "bio_ref.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(f"wrote bio_ref.json: {len(out)} cases, tag2int = {tag2int}")


if __name__ == "__main__":
    main()
