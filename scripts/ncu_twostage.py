"""TwoStageFFTConvolver x 4096 channels, head 512 -> T = 8192, IR 2 s: a short run for ncu (paired head + tail0 launch)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import fft_convolution_b200 as F
C, B, L = 4096, 512, 96000
ts = F.TwoStageFFTConvolver.init(bench.synth_irs(0, C, 0, L), B, L, async_tail=True)
x = torch.from_numpy(bench.synth_noise(0, C, 0, B)).cuda()
out = torch.empty((C, B), dtype=torch.float32, device="cuda")
for _ in range(12):
    ts.process_dev(x.data_ptr(), B, B, out.data_ptr(), B, B)
ts.sync()
print("done")
