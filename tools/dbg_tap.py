"""Phase timestamps of the tcgen05 tap kernel (debug hook)."""
import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from globalegomocap_b200 import synthetic as syn
from globalegomocap_b200.engine import Engine
W = 1870
eng = Engine(max_windows=W)
clip = syn.make_clip(64, seed=3)
sd = syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=syn.mean_pose_bias(clip))
eng.set_vae(0, sd)
z = torch.randn(W, 2048, device="cuda")
for _ in range(3):
    eng.decode(0, z)
torch.cuda.synchronize()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
eng.lib.gem_debug_tap_timestamps.argtypes = [C.c_void_p]
eng.lib.gem_debug_tap_timestamps(C.c_void_p(buf.data_ptr()))
eng.decode(0, z)      # last tap layer (dec[5]) overwrites the buffer last
torch.cuda.synchronize()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(0))
t = buf.view(-1, 16)[:156].cpu().numpy()
names = ["start", "setup done", "kb0 full", "kb1 full", "acc ready", "epilogue done", "end", "chunks done", "ld0", "ld1", "ld2", "ld3"]
base = t[:, 0].min()
for cta in (0, 155):
    print(cta, [(n, int(t[cta, i] - t[cta, 0])) for i, n in enumerate(names)], "start offset", int(t[cta, 0] - base))
d = t[:, 1:12] - t[:, :1]
print("median cycles since CTA start:", dict(zip(names[1:], np.median(d, 0).astype(int))))
print("globaltimer: CTA start spread %d ns, kernel span %d ns, median CTA life %d ns" % (t[:, 14].max() - t[:, 14].min(), t[:, 15].max() - t[:, 14].min(), np.median(t[:, 15] - t[:, 14])))
order = np.argsort(t[:, 14]); print("start offsets (ns) of CTAs sorted:", (t[order, 14] - t[:, 14].min())[[0, 1, 50, 100, 140, 147, 148, 150, 155]])

# same for a split-output layer: run the vjp (its last tap layer, dec_bwd[4], has grid (156, 4))
up = torch.randn(W, 10, 15, 3, device="cuda")
eng.decode_vjp(0, up); torch.cuda.synchronize()
buf.zero_()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(buf.data_ptr()))
eng.decode_vjp(0, up)
torch.cuda.synchronize()
eng.lib.gem_debug_tap_timestamps(C.c_void_p(0))
t = buf.view(-1, 16)[:624].cpu().numpy()
d = t[:, 1:12] - t[:, :1]
print("dec_bwd[4] median cycles since CTA start:", dict(zip(names[1:], np.median(d, 0).astype(int))))
