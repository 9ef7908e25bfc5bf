"""BIO label assignment (reference polus/ner/bio.py:92-114 `get_bio`): entities are visited longest
first; an entity labels tokens only when its span is covered exactly by whole tokens that are still
unlabelled ("O"); the first token gets B-<type>, the rest I-<type>."""


def get_bio(token_spans, entities, default="O"):
    """token_spans: [(start, end)] in text order; entities: [(start, end, type)] -> list of tags."""
    tags = [default] * len(token_spans)
    for es, ee, etype in sorted(entities, key=lambda e: e[1] - e[0], reverse=True):
        idx = [i for i, (ts, te) in enumerate(token_spans) if ts >= es and te <= ee]
        if not idx or token_spans[idx[0]][0] != es or token_spans[idx[-1]][1] != ee:
            continue  # the entity does not fit token boundaries exactly
        if any(tags[i] != default for i in idx):
            continue  # never overwrite a longer entity
        tags[idx[0]] = f"B-{etype}"
        for i in idx[1:]:
            tags[i] = f"I-{etype}"
    return tags
