"""Parity helpers shared by the GPU tests, smoke() and tools/parity_report.py.

Code parity protocol: packed codes are compared frame by frame per utterance.
At the first differing frame t*, every differing bit must lie within EPS of the decision
threshold in the oracle (|sigmoid(logit) - 0.5| < EPS) -> counted as an eps-bit; otherwise it is
a hard mismatch.  The encoder under test is then re-started at t*+1 from the oracle's hidden
state so later frames are still checked.  Masked positions must be exactly 0.5.
"""
from __future__ import annotations

import numpy as np
import torch

EPS_PROB = 1e-4          # |p - 0.5| < 1e-4  <=>  |logit| < 4e-4


def snr_db(ref, test):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return 10.0 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300))


def compare_codes(engine, mel, bits_scalar, oracle_codes, oracle_logits, oracle_all_h, max_resync=64):
    """engine: product _Engine; mel [B,T,X] (torch, on the engine's device).  Returns a report dict."""
    dev = engine.device
    oc = np.asarray(oracle_codes)
    ol = np.asarray(oracle_logits)
    oh = torch.as_tensor(oracle_all_h)
    B, T, Z = oc.shape
    total = int((oc != 0.5).sum())
    rep = dict(total_bits=total, eps_bits=0, hard_mismatches=0, resyncs=0, mask_errors=0, max_logit_err=0.0)
    codes, _, _, logits, _ = engine.encode(mel, None, bits_scalar, None, want_logits=True, want_all_h=False)
    codes = codes.cpu().numpy()
    logits = logits.cpu().numpy()
    for b in range(B):
        start, cur_codes, cur_logits = 0, codes[b], logits[b]
        for _ in range(max_resync + 1):
            seg_o = oc[b, start:]
            diff = cur_codes != seg_o
            if not diff.any():
                rep["max_logit_err"] = max(rep["max_logit_err"], float(np.abs(cur_logits - ol[b, start:]).max()) if len(seg_o) else 0.0)
                break
            t_rel = int(np.argmax(diff.any(axis=1)))
            if t_rel > 0:
                rep["max_logit_err"] = max(rep["max_logit_err"],
                                           float(np.abs(cur_logits[:t_rel] - ol[b, start:start + t_rel]).max()))
            t_abs = start + t_rel
            for i in np.nonzero(diff[t_rel])[0]:
                if seg_o[t_rel, i] == 0.5 or cur_codes[t_rel, i] == 0.5:
                    rep["mask_errors"] += 1
                    continue
                p = 1.0 / (1.0 + np.exp(-float(ol[b, t_abs, i])))
                if abs(p - 0.5) < EPS_PROB:
                    rep["eps_bits"] += 1
                else:
                    rep["hard_mismatches"] += 1
            start = t_abs + 1
            if start >= T:
                break
            rep["resyncs"] += 1
            h0 = oh[b:b + 1, start].to(dev)
            c2, _, _, l2, _ = engine.encode(mel[b:b + 1, start:].contiguous(), None, bits_scalar, h0,
                                            want_logits=True, want_all_h=False)
            cur_codes, cur_logits = c2[0].cpu().numpy(), l2[0].cpu().numpy()
        else:
            rep["hard_mismatches"] += 1   # did not converge within max_resync
    return rep
