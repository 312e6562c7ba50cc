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
swin_ref      torch-fp32 restatement of the reference SwinTransformerV2 (reproduces the class bit for bit)
pca_np        per-model normalize + concat + sklearn PCA.transform (concat_pca_sn.py)
near_dup_np   the near-duplicate frame filter of the query extractor (extract_query_feats.py:190-199)
candidates_np global top-K frame pairs -> video-pair candidates (vsc/index.py, vsc/candidates.py, infer_matching.py)
tn_np         the temporal-network alignment vcsl.vta.tn incl. networkx's dag_longest_path semantics
matching_np   matching-track candidate features (M/infer/src/utils.py, src/dataset.py)
resize_np     Pillow's antialiased bicubic resize (Resample.c) + torchvision ToTensor / Normalize
refload       (container only) imports the reference's own classes from /root/reference
              with import shims; used to pin the restatements and to make tests/golden/*
"""
