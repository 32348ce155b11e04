"""Time the fp32 training step at BASELINE configs[1] shape (1024x436, batch 2) on one GPU."""
import sys, time, torch
sys.path.insert(0, ".")
import vst_b200  # noqa
from vst_b200 import synth
from vst_b200.reconet.network import ReCoNet, Vgg16
from vst_b200.train_core import PairTrainer
H, W, B = 436, 1024, 2
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
if len(sys.argv) > 3 and sys.argv[2].isdigit(): H, W = int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
if "rtnstv" in sys.argv:      # BASELINE configs[2]: 640x360, 4 pairs
    from vst_b200.rtnstv.network import StylizingNetwork
    from vst_b200.rtnstv.vgg19 import VGG19
    H, W, B = 360, 640, 4
    model, vgg, fam = StylizingNetwork().cuda(), VGG19().cuda(), "rtnstv"
    vgg.load_state_dict(synth.vgg_state_dict("vgg19_rt"))
else:
    model, vgg, fam = ReCoNet(1).cuda(), Vgg16().cuda(), "reconet"
    vgg.load_state_dict(synth.vgg_state_dict("vgg16_rc"))
tr = PairTrainer(model, vgg, synth.smooth_frames(1, H, W, "style"), fam, precision=prec)
if "graph" in sys.argv: tr.enable_cuda_graph()
img1, img2 = synth.smooth_frames(B, H, W, "a").cuda(), synth.smooth_frames(B, H, W, "b").cuda()
flow, mask = synth.smooth_flow(B, H, W, "f").cuda(), synth.mask(B, H, W, "m").cuda()
for i in range(2 if "short" in sys.argv else 5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    terms = tr.step(img1, img2, flow, mask)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"step {i}: {dt*1e3:.1f} ms  {B/dt:.2f} pairs/s", terms.to_dict())
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
if "bench" in sys.argv:      # event-timed replay loop (what bench.py's train leg measures)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(40):
        tr.step(img1, img2, flow, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    print(f"BENCH {fam} {prec} {ms:.3f} ms/step  {B / ms * 1e3:.1f} pairs/s")
if "prof" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as pr:
        tr.step(img1, img2, flow, mask)
        torch.cuda.synchronize()
    print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
if "timeline" in sys.argv:   # kernel timeline of ONE (graph-replayed) step: start / duration / stream, per-stream busy time, gaps
    from torch.profiler import ProfilerActivity, profile
    for _ in range(3):
        tr.step(img1, img2, flow, mask)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        tr.step(img1, img2, flow, mask)
        torch.cuda.synchronize()
    ev = [e for e in pr.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    span = max(e.time_range.end for e in ev) - t0
    streams = {}
    for e in ev:
        streams.setdefault(getattr(e, "device_resource_id", getattr(e, "thread", 0)), []).append(e)
    print(f"TIMELINE span {span:.1f} us, {len(ev)} device activities, streams: " +
          ", ".join(f"{k}: {len(v)} kernels busy {sum(x.time_range.end - x.time_range.start for x in v):.0f} us" for k, v in streams.items()))
    # union busy time (any stream) and idle gaps
    pts = sorted([(e.time_range.start - t0, 1) for e in ev] + [(e.time_range.end - t0, -1) for e in ev])
    busy, depth, last = 0.0, 0, 0.0
    for t, d in pts:
        if depth > 0:
            busy += t - last
        depth += d
        last = t
    print(f"TIMELINE any-stream busy {busy:.1f} us, idle {span - busy:.1f} us")
    for e in ev:
        sid = getattr(e, "device_resource_id", getattr(e, "thread", 0))
        print(f"TL {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} s{sid} {e.name[:70]}")
