"""Where does the energy of the layer-4 conv go?  Runs the 512->512 dilation-4 conv (conv_tc2_kernel<256>) back to back for ~2.5 s per mode
under the power cap and reports sustained time per launch, SM clock and power: HK_TC2_DEBUG = 0 (normal), 2 (no TMA operand loads: the
MMAs run on stale shared memory -- no L2->SM operand traffic), 4 (no epilogue body), 1 (no MMAs).  At the cap, time ~ energy / power,
so (t0 - t2) / t0 bounds what removing ALL operand traffic could give; a 4-CTA multicast cluster removes a quarter of it."""
import os, subprocess, sys, threading, time
import torch
# needs the diagnostics build of the library (timeline stamps / HK_TC2_DEBUG switches are compiled out of the shipped .so):
#   python -m hulk_keypoints_b200.build --diag   ->  hulk_keypoints_b200/libhulk_sm100_diag.so
os.environ.setdefault("HK_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hulk_keypoints_b200", "libhulk_sm100_diag.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hulk_keypoints_b200 import ops
dev = torch.device("cuda:0")
B, cin, cout, dil = 64, 512, 512, 4
x = torch.randn(B, 60, 80, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
wp, s, b = ops.pack_conv_weights(w, None, 1e-5, torch.bfloat16)
out = torch.empty(B, 60, 80, cout, device=dev, dtype=torch.bfloat16)
run = lambda: ops.conv_bn_act(x, wp, s, b, stride=1, pad=dil, dil=dil, relu=True, residual=None, out=out)
flops = 2.0 * B * 4800 * cout * cin * 9


def smi():
    r = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True)
    try:
        c, p = r.stdout.strip().split(",")
        return float(c), float(p)
    except Exception:
        return None


for mode in ("0", "2", "0", "2", "4", "1"):
    os.environ["HK_TC2_DEBUG"] = mode
    for _ in range(20):
        run()
    torch.cuda.synchronize()
    samples, stop = [], False

    def sampler():
        while not stop:
            v = smi()
            if v:
                samples.append(v)
            time.sleep(0.2)
    th = threading.Thread(target=sampler); th.start()
    n = 2500
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record(); e1.synchronize()
    stop = True; th.join()
    ms = e0.elapsed_time(e1) / n
    tail = samples[len(samples) // 2:] or [(0, 0)]
    clk = sorted(c for c, _ in tail)[len(tail) // 2]; pw = sorted(p for _, p in tail)[len(tail) // 2]
    print(f"mode={mode}: {ms*1e3:7.1f} us/launch  {flops/ms/1e9:7.1f} TF/s-equivalent  clk {clk:.0f} MHz  power {pw:.0f} W", flush=True)
os.environ["HK_TC2_DEBUG"] = "0"
