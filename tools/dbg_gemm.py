"""Per-CTA phase timestamps of the fp16-scheme GEMM kernels (debug hook), one-CTA vs CTA-pair."""
import ctypes as C, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from globalegomocap_b200 import synthetic as syn
from globalegomocap_b200.engine import Engine
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1870
eng = Engine(max_windows=W)
clip = syn.make_clip(64, seed=3)
sd = syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=syn.mean_pose_bias(clip))
eng.set_vae(0, sd)
eng.set_gemm_mode(3)
lib = eng.lib
lib.gem_debug_gemm_timestamps.argtypes = [C.c_void_p]
lib.gem_debug_gemm_pair.argtypes = [C.c_int]
z = torch.randn(W, 2048, device="cuda")
up = torch.randn(W, 10, 15, 3, device="cuda")
names = ["start", "setup", "producer done", "kb0 full", "mma issued", "acc ready", "epi done", "end", "wait_empty", "wait_full"]
buf = torch.zeros(16 * 4096, dtype=torch.int64, device="cuda")
for pair in (0, 1):
    lib.gem_debug_gemm_pair(pair)
    for what in ("fwd latent->T*256", "bwd T*256->latent"):
        for _ in range(2):
            eng.decode(0, z); eng.decode_vjp(0, up)
        torch.cuda.synchronize()
        buf.zero_()
        if what.startswith("fwd"):
            lib.gem_debug_gemm_timestamps(C.c_void_p(buf.data_ptr()))
            eng.decode(0, z)
        else:
            eng.decode(0, z)
            lib.gem_debug_gemm_timestamps(C.c_void_p(buf.data_ptr()))
            eng.decode_vjp(0, up)
        torch.cuda.synchronize()
        lib.gem_debug_gemm_timestamps(C.c_void_p(0))
        t = buf.view(-1, 16).cpu().numpy()
        t = t[t[:, 0] != 0]
        d = t[:, 1:8] - t[:, :1]
        lead = t[t[:, 9] != 0]
        print(f"pair={pair} {what}: {len(t)} CTAs; median cycles since CTA start:",
              {n: int(v) for n, v in zip(names[1:8], np.median(d, 0))})
        print("   leader/MMA CTAs: median wait_full %d, producer wait_empty %d (all CTAs), mma issued - kb0 full %d" % (
            np.median(lead[:, 9]), np.median(t[:, 8]), np.median(lead[:, 4] - lead[:, 3])))
        print("   globaltimer: kernel span %d ns, median CTA life %d ns, CTA start spread %d ns" % (
            t[:, 15].max() - t[:, 14].min(), np.median(t[:, 15] - t[:, 14]), t[:, 14].max() - t[:, 14].min()))
        order = np.argsort(t[:, 14])
        print("   start offsets (ns) sorted [0,100,147,148,200,-1]:", [(int(t[order[i], 14] - t[:, 14].min())) for i in (0, min(100, len(t) - 1), min(147, len(t) - 1), min(148, len(t) - 1), min(200, len(t) - 1), -1)])
lib.gem_debug_gemm_pair(-1)
