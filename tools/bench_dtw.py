"""DTW throughput (cells/s) of wca_dtw_align on BASELINE-shaped batches, next to the single-core C oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi
from oracle import dtw as odtw

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def run(label, n_prob, N, M, reps=5):
    costs = torch.from_numpy(-np.abs(rng.standard_normal((n_prob, N, M)).astype(np.float32))).to(dev)
    recs = np.zeros(n_prob, dtype=_cabi.UTT_DTYPE)
    for b in range(n_prob):
        recs[b]["n_tokens"], recs[b]["n_frames"], recs[b]["row_begin"], recs[b]["row_end"] = N, M, 0, N
        recs[b]["matrix_off"], recs[b]["jump_off"] = b * N * M, b * N
    d_utts = _cabi.upload_utts(recs, dev)
    jumps = torch.empty(n_prob * N, dtype=torch.int32, device=dev)
    nbytes = _cabi.dtw_workspace_bytes(n_prob, N, M)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None
    for _ in range(2):
        _cabi.dtw_align(costs.data_ptr(), d_utts, n_prob, N, M, False, jump_frames=jumps, trace_ws=ws)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _cabi.dtw_align(costs.data_ptr(), d_utts, n_prob, N, M, False, jump_frames=jumps, trace_ws=ws)
    b_.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / reps
    # oracle check + CPU rate on one problem
    c0 = costs[0].cpu().numpy()
    t0 = time.perf_counter(); ti, tj = odtw.dtw_path(c0); cpu_s = time.perf_counter() - t0
    want = odtw.jump_frames(ti, tj)
    assert np.array_equal(jumps[:N].cpu().numpy(), want), "jump frames differ from the oracle"
    cells = n_prob * N * M
    print(f"{label:34s} {n_prob:5d} x ({N:3d} x {M:4d}): {ms*1e3:9.1f} us/launch  {cells/ms/1e6:9.2f} Gcells/s   "
          f"(C oracle, 1 core: {N*M/cpu_s/1e6:6.1f} Mcells/s; trace ws {nbytes} B)", flush=True)

run("config 2, one batch of 16", 16, 41, 150)
run("config 2, 1680 utterances", 1680, 41, 150)
run("config 5, 384 heads x 16 utts", 6144, 41, 150)
run("config 5 probe-shaped (>=18 words)", 384 * 4, 96, 300)
run("config 3, batch of 8", 8, 401, 1500, reps=3)
run("config 3, 64 utterances", 64, 401, 1500, reps=2)
run("largest legal problem", 4, 445, 1500, reps=2)
