"""Times the fp16-scheme GEMM alone (split kernel + GEMM per call) in its one-CTA and CTA-pair variants."""
import ctypes as C, os, sys, json
import torch
sys.path.insert(0, os.getcwd())
from globalegomocap_b200.engine import Engine
eng = Engine(max_windows=64)
lib = eng.lib
lib.gem_debug_gemm_pair.argtypes = [C.c_int]
res = []
for (M, N, K) in [(1870, 2560, 2048), (1870, 2048, 2560), (468, 2560, 2048), (468, 2048, 2560)]:
    a = torch.randn(M, K, device="cuda"); b = torch.randn(K, N, device="cuda") / K ** 0.5
    for pair, bn in [(0, 0), (0, 128), (0, 160), (1, 160), (1, 128)]:
        lib.gem_debug_gemm_pair(pair)
        if bn: os.environ["GEM_GEMM_BN"] = str(bn)
        else: os.environ.pop("GEM_GEMM_BN", None)
        for _ in range(3): eng.gemm(a, b, None, tensor_cores=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        with torch.cuda.stream(torch.cuda.ExternalStream(eng.stream) if isinstance(eng.stream, int) and eng.stream else torch.cuda.current_stream()):
            e0.record()
            for _ in range(n): eng.gemm(a, b, None, tensor_cores=2)
            e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        r = dict(M=M, N=N, K=K, pair=pair, bn=bn, us=round(us, 2), tflops_fp32eq=round(2.0 * M * N * K / us / 1e6, 1))
        print(json.dumps(r), flush=True)
        res.append(r)
