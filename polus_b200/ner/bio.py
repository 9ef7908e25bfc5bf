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


def decode_bio(tags, spans, text=None, allow_errors=False):
    """Inverse of get_bio (reference polus/ner/bio.py:115-188 `decode_bio`): BIO tags + token spans -> the entity set
    {(start, end, type)} and the decoding statistics the reference logs.

    Semantics kept from the reference: an entity is closed by "O", by the next "B-*", by an "I-*" of another type, or by
    the end of the sequence; with allow_errors an "I-*" after "O" opens an entity (counted in
    `inside_tag_after_other_tag`) and an "I-*" of a different type closes the running entity and opens a new one
    (`inside_tag_with_different_entity_type`); without it both cases assert.  One deliberate difference: the reference
    re-initialises the running (start, end, type) inside its loop at HEAD, so every entity it emits is (-1, -1, "None")
     -- here the state is carried across tokens, which is what its comments and its callers expect.
    `text` is accepted for signature compatibility; entities are identified by (span, type)."""
    entities = set()
    counts = {"tags": len(tags), "inside_tag_after_other_tag": 0, "inside_tag_with_different_entity_type": 0}
    s = e = -1
    t = None
    open_ = False
    for tag, (start, end) in zip(tags, spans):
        if tag in ("O", "PAD"):
            if open_:
                entities.add((s, e, t))
            open_ = False
        elif tag[0] == "B":
            if open_:
                entities.add((s, e, t))
            s, e, t, open_ = start, end, tag.split("-", 1)[1], True
        elif tag[0] == "I":
            t_new = tag.split("-", 1)[1]
            if not open_:
                assert allow_errors, "Previous token does not belong to an Entity."
                s, t, open_ = start, t_new, True
                counts["inside_tag_after_other_tag"] += 1
            elif t_new != t:
                assert allow_errors, "Previous Entity type is different."
                entities.add((s, e, t))
                s, t = start, t_new
                counts["inside_tag_with_different_entity_type"] += 1
            e = end
        else:
            raise ValueError(f"not a BIO tag: {tag!r}")
    if open_:
        entities.add((s, e, t))
    return entities, counts
