"""Host-side text helpers the drop-in classes need around the hot path.

These are NOT accelerated (chunk-text reconstruction and tokenisation stay
on the host); they exist so the mirror classes return the same dict payloads as the
reference.  Behaviour follows rag_search_engine/utils/utils.py:126-179 (``chunk`` /
``semantic_chunk``) and is checked against the reference's functions in
tests/test_host_cpu.py when the reference checkout is present.
"""
from __future__ import annotations

import re
from typing import Callable, List, Sequence

_SENTENCE_SPLIT = re.compile(r"(?<=[.!?])\s+")       # utils.py:169


def window_chunks(items: Sequence[str], chunk_size: int, overlap: int) -> List[List[str]]:
    """utils.py:126-155 as window arithmetic.

    The first window is ``items[:chunk_size]``.  Every further window starts with the last
    ``overlap`` items of the previous FULL window (only when 0 < overlap < chunk_size) and is
    filled with new items up to ``chunk_size``; a trailing partial window is kept.
    """
    items = list(items)
    out = [items[:chunk_size]]
    carry = overlap if (0 < overlap < chunk_size) else 0
    fresh = chunk_size - carry                      # new items per later window
    pos = chunk_size
    n = len(items)
    while pos < n:
        head = out[-1][-carry:] if carry else []
        out.append(list(head) + items[pos:pos + fresh])
        pos += fresh
    return out


def sentence_chunks(text: str, max_chunk_size: int, overlap: int) -> List[List[str]]:
    """utils.py:158-179 for a single string: split on sentence punctuation, then window."""
    return window_chunks(_SENTENCE_SPLIT.split(text.strip()), max_chunk_size, overlap)


def chunk_text(title: str, description: str, chunk_index: int, max_chunk_size: int, overlap: int) -> str:
    """semantic_search.py:342-366: chunk 0 is the title, chunk i ≥ 1 the (i-1)-th description
    window joined WITHOUT a separator (:182); out-of-range falls back to the description."""
    if chunk_index == 0:
        return title
    chunks = sentence_chunks(description, max_chunk_size, overlap)
    i = chunk_index - 1
    if i < 0 or i >= len(chunks):
        return description
    return "".join(chunks[i])


Tokenizer = Callable[[List[str]], List[List[str]]]


def whitespace_tokenizer(texts) -> List[List[str]]:
    """Frozen stand-in for the reference's spaCy ``preprocess`` (utils.py:75-123): lower-case and
    split on whitespace.  Used for the synthetic corpora and the tests."""
    if isinstance(texts, str):
        texts = [texts]
    return [t.lower().split() for t in texts]


def default_tokenizer() -> Tokenizer:
    """The reference's own ``preprocess`` when the reference package (and spaCy) is installed."""
    try:
        from rag_search_engine.utils.utils import preprocess  # type: ignore
    except Exception as e:  # pragma: no cover - depends on the environment
        raise RuntimeError(
            "no tokenizer available: the reference's spaCy `preprocess` "
            "(rag_search_engine.utils.utils) is not importable here. Pass tokenizer=... "
            "(e.g. rag_search_engine_b200.textutil.whitespace_tokenizer for frozen corpora) "
            "or use the *_tokens entry points.") from e
    return preprocess
