"""Label contract of polus.ner (reference polus/ner/utils.py:9-15)."""
TAG2INT = {"PAD": 0, "O": 1, "B-Chemical": 2, "I-Chemical": 3}
INT2TAG = {v: k for k, v in TAG2INT.items()}
