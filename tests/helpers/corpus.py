"""Shared corpus helpers for the tests (frozen whitespace tokenizer, reference index layout)."""
import numpy as np

TITLE_END = "[TITLE_END]"


def build_postings(docs):
    """keyword_search.py:102-177 with the frozen whitespace tokenizer → {term: [(doc, tf)]} asc doc."""
    postings, doclen = {}, {}
    for d in docs:
        toks = d["title"].lower().split() + [TITLE_END] + d["description"].lower().split()
        doclen[int(d["id"])] = len(toks)
        cnt = {}
        for t in toks:
            cnt[t] = cnt.get(t, 0) + 1
        for t, c in cnt.items():
            postings.setdefault(t, []).append((int(d["id"]), c))
    for t in postings:
        postings[t].sort()
    return postings, doclen


def to_csr(postings, doclen):
    doc_ids = np.array(sorted(doclen), np.int64)
    dense = {int(d): i for i, d in enumerate(doc_ids)}
    terms = sorted(postings)
    row = {t: i for i, t in enumerate(terms)}
    indptr = [0]
    doc, tf = [], []
    for t in terms:
        for d, c in postings[t]:
            doc.append(dense[d]); tf.append(c)
        indptr.append(len(doc))
    dl = np.array([doclen[int(d)] for d in doc_ids], np.uint32)
    df = np.diff(np.array(indptr, np.int64))
    return dict(indptr=np.array(indptr, np.int64), doc=np.array(doc, np.uint32), tf=np.array(tf, np.uint32),
                df=df, dl=dl, doc_ids=doc_ids, row=row, avgdl=float(dl.sum()) / len(dl))


