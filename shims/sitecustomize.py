"""Picked up automatically by the interpreter when <repo>/shims is on PYTHONPATH.  With VSCB200_ACTIVATE=1 the unmodified
reference scripts (D/infer/eval.sh, infer_ref.sh, infer_query.sh; SURVEY.md 8b seam C) run their hot path on the B200
library: see vsc22_submission_b200/activate.py.  Without the variable this file does nothing."""
import os

if os.environ.get("VSCB200_ACTIVATE", "") not in ("", "0"):
    from vsc22_submission_b200.activate import activate
    activate(int(os.environ.get("VSCB200_MAX_FRAMES", "256")))
