"""`eval_ali.py` shell: re-score a `*-predictions.pkl` at another tolerance (reference eval_ali.py:9-53).
Pure host code."""
from __future__ import annotations

import argparse

import joblib

from ..metrics import eval_n1_strict, get_seg_metrics
from ..retokenize import remove_punctuation


def run_eval(args):
    preds = joblib.load(args.pred)
    corrects = total_preds = total_gts = 0
    for i in sorted(k for k in preds if preds[k]):  # skipped utterances were stored as 0
        p = preds[i]
        ref_words = [remove_punctuation(w) for w in p["texts"]]
        hyp_words = [remove_punctuation(w) for w in p["predwords"]]
        tp, fp, fn = eval_n1_strict(p["ends"], p["ends_hat"], ref_words, hyp_words, tolerance=args.tolerance)
        corrects += tp
        total_gts += tp + fn
        total_preds += tp + fp
    precision, recall, f1, r_value, _ = get_seg_metrics(corrects, corrects, total_preds, total_gts)
    print("-----------------")
    print(f"precision: {precision:.2f}")
    print(f"recall: {recall:.2f}")
    print(f"f1: {f1:.2f}")
    print(f"r value: {r_value:.2f}")
    print("-----------------")
    return dict(precision=precision, recall=recall, f1=f1, r_value=r_value)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="eval alignment")
    p.add_argument("--pred", type=str, required=True)  # /path/to/*-predictions.pkl
    p.add_argument("--tolerance", type=float, default=0.05)
    return p.parse_args(argv)


def main(argv=None):
    return run_eval(parse_args(argv))


if __name__ == "__main__":
    main()
