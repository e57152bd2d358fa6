"""Golden vectors for the selected-sentence-index output, FROM THE REFERENCE ITSELF.

Imports the reference's own ``evaluate.py`` (unmodified, /root/reference or the staged oracle/_ref copy) and runs its
``get_generated_summaries`` -> ``greedy_search`` -> ``get_source_sentence`` (evaluate.py:167-202, :236-259) on seeded
distributions, with real pickled "source sentence" files on disk in the reference's format (a list of (sentence, ...) tuples,
preprocess_text.py:78-85).  Modules evaluate.py imports but the hot path never uses (seaborn, rouge, tensorboardX, ujson, nltk,
tqdm, sklearn, PIL, torchvision, datasets, models, util, args) are stubbed; ``datasets`` is what gives evaluate.py its ``pickle``
and ``logging`` names through ``from datasets import *``, so the stub carries exactly those.

    python tests/golden/make_golden_greedy.py            # rewrites tests/golden/greedy_search.pt
"""
import importlib.util
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MMBIDAF_REFERENCE", "/root/reference")
if not os.path.exists(os.path.join(REF, "evaluate.py")):
    REF = os.path.join(REPO, "oracle", "_ref")


def import_reference_evaluate():
    import logging
    stubs = {}

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        stubs[name] = mod
        return mod
    anything = type("Anything", (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
    stub("seaborn")
    stub("rouge", Rouge=anything)
    stub("tensorboardX", SummaryWriter=anything)
    stub("tqdm", tqdm=lambda x, *a, **k: x)
    stub("ujson", load=lambda *a, **k: None)
    stub("nltk", PorterStemmer=anything)
    stub("nltk.tokenize", sent_tokenize=lambda s: [s])
    stub("sklearn")
    stub("sklearn.metrics", f1_score=lambda *a, **k: 0.0)
    stub("PIL", Image=anything)
    stub("torchvision")
    stub("torchvision.transforms")
    stub("datasets", pickle=pickle, logging=logging, __all__=["pickle", "logging"])
    stub("models", MMBiDAF=anything)
    stub("util")
    stub("args", get_train_args=lambda: None)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("reference_evaluate", os.path.join(REF, "evaluate.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    ev = import_reference_evaluate()
    gen = torch.Generator().manual_seed(77)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        def source_file(n_sentences, tag):
            path = os.path.join(tmp, f"src_{tag}.pkl")
            with open(path, "wb") as f:
                pickle.dump([(f"sentence {i}", f"raw {i}") for i in range(n_sentences)], f)
            return path
        # (B, T, M) batches with ragged text lengths; text length INCLUDES the EOS row (datasets.py:70), so a transcript with
        # n sentences has original_text_length n + 1 and EOS index n
        for case, (bsz, t_dec, m) in enumerate(((4, 7, 12), (3, 12, 40), (5, 5, 9))):
            dist = torch.rand(bsz, t_dec, m, generator=gen)
            lengths = torch.randint(2, m + 1, (bsz,), generator=gen).tolist()
            for b, n in enumerate(lengths):
                dist[b, :, n:] = 0.0                                   # the decoder mask zeroes padded sentences (models.py:119-123)
            if case == 0:
                dist[0, 0, lengths[0] - 1] = 5.0                        # EOS at the first step: empty summary
                dist[1, 2, lengths[1] - 1] = 5.0                        # EOS at step 2
                if lengths[2] < m:
                    dist[2, 1, m - 1] = 9.0                             # an index beyond the transcript: skipped, not a stop (evaluate.py:256-257)
            if case == 2:
                dist[3, :, 0] = 7.0                                     # the same sentence picked every step (no coverage zeroing, :199-200)
            paths = [source_file(n - 1, f"{case}_{b}") for b, n in enumerate(lengths)]
            _, gen_idxs = ev.get_generated_summaries(dist, lengths, paths, method="greedy")
            want = [[int(i) for i in idxs[0]] for idxs in gen_idxs]
            cases.append({"dist": dist, "lengths": lengths, "indices": want})
    torch.save({"cases": cases, "numpy": np.__version__, "reference": "evaluate.py:167-202, :236-259"},
               os.path.join(HERE, "greedy_search.pt"))
    for c in cases:
        print(c["lengths"], c["indices"])


if __name__ == "__main__":
    main()
