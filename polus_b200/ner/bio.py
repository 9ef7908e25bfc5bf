"""BIO label assignment (reference polus/ner/bio.py:92-114 `get_bio`): entities are visited longest
first; an entity labels tokens only when its span is covered exactly by whole tokens that are still
unlabelled ("O"); the first token gets B-<type>, the rest I-<type>."""


def _entity_fields(entity):
    """(start, end, type) of a tuple entity or of an object with .start / .end / .typ (the reference's Entity)."""
    if isinstance(entity, (tuple, list)):
        return entity[0], entity[1], entity[2]
    return entity.start, entity.end, entity.typ


def longer_entities_first(entities):
    """Longer spans first: they win when entities overlap (reference bio.py:5-10; stable for equal lengths)."""
    return sorted(entities, key=lambda e: _entity_fields(e)[1] - _entity_fields(e)[0], reverse=True)


def entity_fits_spans(entity, spans):
    """Indexes of the consecutive token spans that cover the entity EXACTLY (first span starts at entity.start, last one
    ends at entity.end), or False (reference bio.py:13-43).  `spans` are in text order."""
    start, end, _ = _entity_fields(entity)
    indexes = []
    for i, (s0, s1) in enumerate(spans):
        if not indexes:
            if s0 > start:
                return False  # walked past the entity's first character without a token starting there
            if s0 != start:
                continue
        indexes.append(i)
        if s1 == end:
            return indexes
        if s1 > end:
            return False
    return False


def update_tags(tags, spans, entity, default="O"):
    """Writes B-<type> / I-<type> over the entity's tokens, only if it fits token boundaries exactly and none of its tokens
    is already labelled (reference bio.py:46-89).  In place."""
    idx = entity_fits_spans(entity, spans)
    if idx and all(tags[i] == default for i in idx):
        etype = _entity_fields(entity)[2]
        tags[idx[0]] = f"B-{etype}"
        for i in idx[1:]:
            tags[i] = f"I-{etype}"


def get_bio(token_spans, entities, default="O"):
    """token_spans: [(start, end)] in text order; entities: [(start, end, type)] -> list of tags (reference bio.py:92-114)."""
    tags = [default] * len(token_spans)
    for entity in longer_entities_first(entities):
        update_tags(tags, token_spans, entity, default)
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
