"""CPU oracle for the VSC22 hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product package
(``vsc22_submission_b200``) must not import from here and fails loudly when its
CUDA library is missing.

Modules
-------
faiss_np      exact numpy restatement of the faiss ``IndexFlat`` subset the reference calls
score_norm_np restatement of ``vsc/baseline/score_normalization.py`` on plain arrays
vit_ref       torch-fp32 CPU restatement of the reference ViT encoders (+GeM tails)
refload       (container only) imports the reference's own classes from /root/reference
              with import shims; used to pin the restatements and to make tests/golden/*
"""
