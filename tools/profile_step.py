"""One eager SNIPER training step (B = 20 chips, config 2 or --bf16) bracketed by cudaProfilerStart/Stop after one
warm-up step: the command ncu wraps for the per-kernel captures under profiles/ (B200_PROFILING.md recipe), e.g.
  ncu --set full --clock-control none --profile-from-start off -k regex:deform_psroi -c 4 -o gpurun_out/ncu_psroi \
      python tools/profile_step.py
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sniper_b200 import model, synth_batch  # noqa: E402


def main():
    cfg = model.Cfg()
    cfg.batch_images = int(os.environ.get("CHIPS", "20"))
    cfg.bf16 = "--bf16" in sys.argv
    cfg.wgrad_stream = False           # one stream: launches appear in program order
    net = model.SniperResNet101(cfg, deform_offset_std=0.01)
    batch = synth_batch.make_batch(cfg.batch_images, seed=100, device="cuda")
    net.train_step(batch, lr=0.0005)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net.train_step(batch, lr=0.0005)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
