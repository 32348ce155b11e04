"""X9 prologue A/B: the same fp32 frames through vst_tc_prologue_x9 with VST_X9_STAGED=2 (default: source segment staged in
shared memory as 16-bit RGB + rows leaving through a swizzled tile as 512-byte warp stores), =1 (staged store only) and =0
(every thread gathers its 27 elements through L1 and stores its own 64-byte row); prints us per launch (L2 flushed between
launches), GB/s over the algorithmic bytes and a digest of the operand - the digests must be equal (also on ragged widths)."""
import hashlib, json, os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import vst_b200  # noqa
    from vst_b200 import tc
    out = {}
    for (N, H, W) in ((4, 1080, 1920), (2, 100, 333), (1, 37, 45), (2, 64, 257), (1, 8, 5)):
        g = torch.Generator("cuda").manual_seed(N * 1000 + W)
        x = torch.rand((N, 3, H, W), device="cuda", generator=g) * 255
        a = tc.prologue_x9(x, 32)
        torch.cuda.synchronize()
        raw = a.t.view(torch.int16).cpu().numpy().tobytes()
        key = f"{N}x{H}x{W}"
        out[key] = {"digest": hashlib.sha1(raw).hexdigest()}
        if H == 1080:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ts = []
            for _ in range(20):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); tc.prologue_x9(x, 32); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            by = x.numel() * 4 + N * (H + 8) * W * 64
            out[key].update(us_median=ts[len(ts) // 2], us_min=ts[0], gbs=by / ts[len(ts) // 2] / 1e3)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        res = {}
        names = {"2": "segment", "1": "staged_store", "0": "direct"}
        for sg in ("2", "1", "0"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, VST_X9_STAGED=sg),
                               capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stderr[-2000:]); sys.exit(1)
            res[names[sg]] = json.loads(r.stdout.strip().splitlines()[-1])
        res["bit_identical"] = all(res[m][k]["digest"] == res["direct"][k]["digest"] for m in ("segment", "staged_store")
                                   for k in res["direct"])
        print(json.dumps(res))
        sys.exit(0 if res["bit_identical"] else 2)
