"""Phase timestamps of the tcgen05 tap-chain kernel (debug hook): fwd chain (dec[2..5]) and bwd chain (dec_bwd[0..3])."""
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
z = torch.randn(W, 2048, device="cuda")
up = torch.randn(W, 10, 15, 3, device="cuda")
for _ in range(3):
    eng.decode(0, z); eng.decode_vjp(0, up)
torch.cuda.synchronize()
nct = (W + 11) // 12
buf = torch.zeros(16 * 4096, dtype=torch.int64, device="cuda")
eng.lib.gem_debug_tap_timestamps.argtypes = [C.c_void_p]
names = {1: "setup", 2: "mma L0", 3: "mma L1", 4: "mma L2", 5: "epi done", 6: "end", 8: "acc L0", 9: "acc L1", 10: "acc L2", 11: "acc L3", 12: "acc L4", 13: "acc L5"}
nct = 2 * ((nct + 1) // 2) if True else nct
def report(tag):
    t = buf.view(-1, 16)[:nct].cpu().numpy()
    d = {n: int(np.median(t[:, i] - t[:, 0])) for i, n in sorted(names.items())}
    print(tag, "median cycles since CTA start:", d)
    print("   globaltimer: kernel span %d ns, median CTA life %d ns, start spread %d ns" % (
        t[:, 15].max() - t[:, 14].min(), np.median(t[:, 15] - t[:, 14]), t[:, 14].max() - t[:, 14].min()))
buf.zero_()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(buf.data_ptr()))
eng.decode(0, z)          # tap 101 writes first, the fwd chain overwrites the buffer last
torch.cuda.synchronize()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(0))
report("fwd chain")
eng.decode(0, z); torch.cuda.synchronize()
buf.zero_()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(buf.data_ptr()))
eng.decode_vjp(0, up)
torch.cuda.synchronize()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(0))
report("bwd chain")
