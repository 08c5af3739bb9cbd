"""Command-line shells with the reference's flags (infer_ali.py, probe_oracle.py, eval_ali.py)."""
