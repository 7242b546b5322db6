import sys
import torch
for ph in ("fwd", "bwd", "grad"):
    a, b = torch.load(f"/tmp/a_{ph}.pt"), torch.load(f"/tmp/b_{ph}.pt")
    a, b = a.flatten(), b.flatten()
    d = (a - b).abs()
    d[torch.isnan(d)] = 1e30
    scale = torch.maximum(a.abs(), b.abs()).clamp_min(1e-20)
    bad = ((d / scale > 1e-3) & (d > 1e-7)).nonzero().flatten()
    print(ph, "numel", a.numel(), "differing (rel>1e-3)", bad.numel())
    if bad.numel():
        # contiguous ranges (gaps > 65536 floats start a new range)
        idx = bad.tolist()
        start = prev = idx[0]; cnt = 1
        for i in idx[1:]:
            if i - prev > 65536:
                print("   range floats [%d, %d] bytes [%d, %d] count %d maxdiff %.3e" % (start, prev, start * 4, prev * 4 + 3, cnt, float(d[start:prev + 1].max())))
                start = i; cnt = 0
            prev = i; cnt += 1
        print("   range floats [%d, %d] bytes [%d, %d] count %d maxdiff %.3e" % (start, prev, start * 4, prev * 4 + 3, cnt, float(d[start:prev + 1].max())))
