#!/usr/bin/env python
"""The UNMODIFIED reference's command line (`optimize_whole_sequence.py`, reference file of the same name) on a small
synthetic dataset of three clips, run in this CPU container: its printed summary is the golden
(tests/golden/cli_summary.json).  The script is executed as `__main__` through runpy from the harness's scratch CWD
(open3d / natsort stubbed, random-init checkpoints at the hard-coded paths); the only thing set from outside is
`torch.manual_seed(SEED)` right before, which is what `--seed` does in this repository's CLI.

    python tests/golden/make_golden_cli.py [--threads 8]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import re
import runpy
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

syn = mg.syn
SEED = 0
CLIPS = (("data_start_0_end_26", 26, 61), ("data_start_26_end_60", 34, 62), ("data_start_100_end_118", 18, 63))   # name, frames, seed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=8)
    args = ap.parse_args()
    import torch
    torch.set_num_threads(args.threads)
    h = mg.Harness(tempfile.mkdtemp(prefix="gem_golden_cli_"), syn.make_clip(58, seed=7))
    # the harness stubs the absent `natsort` package with plain `sorted`; this script's clip names need the real
    # natural order (…_0_…, …_26_…, …_100_…) that natsort.natsorted gives (optimize_whole_sequence.py:48)
    sys.modules["natsort"].natsorted = lambda names: sorted(
        names, key=lambda t: [int(u) if u.isdigit() else u.lower() for u in re.split(r"(\d+)", t)])
    data_root = os.path.join(h.scratch, "data", "synth_cli")
    for name, frames, seed in CLIPS:
        syn.write_clip_pickle(syn.make_clip(frames, seed=seed), os.path.join(data_root, name))
    argv = sys.argv
    sys.argv = ["optimize_whole_sequence.py", "--data_path", "data/synth_cli", "--camera", h.camera_json]
    buf = io.StringIO()
    torch.manual_seed(SEED)
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(mg.REF, "optimize_whole_sequence.py"), run_name="__main__")
    finally:
        sys.argv = argv
    text = buf.getvalue()
    lines = [ln for ln in text.splitlines() if ln.startswith(("Average", "running data", "-----"))]
    summary = {}
    for ln in lines:
        m = re.match(r"(Average [^:]+): (.*)$", ln)
        if m:
            summary[m.group(1)] = float(m.group(2))
    m = re.search(r"joints error is: \[([^\]]*)\]", text)          # a 15-vector printed over several lines
    joints_error = [float(t) for t in m.group(1).split()]
    out = {"seed": SEED, "threads": args.threads, "clips": [list(c) for c in CLIPS], "lines": lines, "summary": summary,
           "joints_error": joints_error}
    with open(os.path.join(mg.OUT, "cli_summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
