"""Run one constrained k-means configuration per process (a CUDA fault is sticky): python tools/kmeans_debug.py <case>"""
import importlib, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
CASES = {"unc_small": ([3000], [3], 0, 0), "min_small": ([9000], [4], 2048, 0), "bal_small": ([3 * 2048], [3], 2048, 2048),
         "bal_k1": ([2048], [1], 2048, 2048), "bal_big": ([9 * 2048], [9], 2048, 2048), "unc_big": ([20000], [5], 0, 0)}
if len(sys.argv) < 2:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True)
        print(c, "rc", r.returncode, (r.stdout + r.stderr).strip().splitlines()[-1:] )
    sys.exit(0)
import torch
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
from oracle import kmeans_oracle as ko
sizes, ks, smin, smax = CASES[sys.argv[1]]
x = np.random.default_rng(1).random((sum(sizes), 3), dtype=np.float32)
off = np.concatenate([[0], np.cumsum(sizes)])
lab, cent, it = amp.kmeans_constrained_windows(torch.from_numpy(x).cuda(), off, ks, smin, smax)
torch.cuda.synchronize()
el, ec, eit = ko.kmeans_constrained(x, ks[0], smin or None, smax or None)
print("ok", bool((lab.cpu().numpy() == el).all()), bool((cent[0, :ks[0]].cpu().numpy() == ec).all()), int(it[0]), eit)
