"""4096 channels x 2 s IR x block 512, 4 blocks per call (time-batched pass) — a short run for ncu."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F
C, B, L, nb = 4096, 512, 96000, 4
conv = F.FFTConvolver.init(bench.synth_irs(0, C, 0, L), B, L)
x = torch.from_numpy(bench.synth_noise(0, C, 0, B * nb)).cuda()
out = torch.empty((C, B * nb), dtype=torch.float32, device="cuda")
for _ in range(6):
    conv.process_dev(x.data_ptr(), B * nb, B * nb, out.data_ptr(), B * nb, B * nb)
conv.sync() if hasattr(conv, "sync") else torch.cuda.synchronize()
torch.cuda.synchronize()
print("done")
