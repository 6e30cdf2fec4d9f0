"""Phase timeline of fps_cluster_kernel (clock64 stamps of three warps of CTA 0 over picks 100..107).

    python tools/fps_prof.py build      # here: nvcc -DAMP_FPS_PROF fps.cu abi_common.cu -> tools/_build/libfpsprof.so
    python tools/fps_prof.py            # on the GPU box
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "3d-semantic-segmentation-amp-net_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_build", "libfpsprof.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-DAMP_FPS_PROF", "-shared", "-o", OUT, os.path.join(CSRC, "fps.cu"), os.path.join(CSRC, "abi_common.cu")])
    print(OUT); sys.exit(0)
import numpy as np, torch
lib = ctypes.CDLL(OUT)
B, P, S = 64, 40000, 2048
pc = torch.rand(B, P, 11, device="cuda")
idx = torch.empty(B, S, dtype=torch.int64, device="cuda")
st = torch.zeros(B, dtype=torch.int32, device="cuda")
vp = ctypes.c_void_p
lib.amp_fps_f32.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, ctypes.c_size_t, vp]
for _ in range(2):
    rc = lib.amp_fps_f32(pc.data_ptr(), B, P, 11, S, 0, idx.data_ptr(), st.data_ptr(), None, 0, None)
    torch.cuda.synchronize()
assert rc == 0
buf = (ctypes.c_longlong * (3 * 8 * 6))()
lib.amp_fps_prof_dump(buf)
a = np.array(buf[:]).reshape(3, 8, 6)
names = ["loop", "max+rescan", "post", "wait", "reduce"]
for w, wn in enumerate(("warp 0", "warp 15", "warp 31")):
    print(wn)
    for s in range(8):
        t = a[w, s]
        nxt = a[w, s + 1, 0] - t[5] if s < 7 else 0
        print("  pick %d: " % (100 + s) + "  ".join("%s %5d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)) + "   total %5d  (to next %d)" % (t[5] - t[0], nxt))
print("pick period (warp 0):", [int(a[0, s + 1, 0] - a[0, s, 0]) for s in range(7)])
print("warp 15 - warp 0 at loop start:", [int(a[1, s, 0] - a[0, s, 0]) for s in range(8)], " at loop end:", [int(a[1, s, 1] - a[0, s, 1]) for s in range(8)])
print("warp 31 - warp 0 at loop start:", [int(a[2, s, 0] - a[0, s, 0]) for s in range(8)], " at loop end:", [int(a[2, s, 1] - a[0, s, 1]) for s in range(8)])
