"""Headline step (configs[2], batch 64) under torch.profiler: sum of the device time of every kernel of one step against the
step's event-timed duration (the difference is launch gaps / host stalls), and the per-kernel totals without ncu's
serialisation."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
dev = torch.device("cuda", 0)
agent = bench.make_agent(bench.cfg_of("cfg3"), dev)
x = S.synthetic_rgb(64, 512, 768, 1337).to(dev)
ms = bench.ev_time(lambda: agent.validate_batch_async(x), 3, 2)
print(f"event-timed step: {ms:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    agent.validate_batch_async(x)
    torch.cuda.synchronize()
agg, cnt = collections.Counter(), collections.Counter()
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:70]
        agg[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        cnt[name] += 1
        s, e = ev.time_range.start, ev.time_range.end
        t0 = s if t0 is None else min(t0, s)
        t1 = e if t1 is None else max(t1, e)
tot = sum(agg.values())
print(f"device span of the profiled step: {(t1 - t0) / 1e3:.1f} ms, sum of kernel / memcpy device time: {tot / 1e3:.1f} ms, "
      f"{sum(cnt.values())} device activities")
for n, v in agg.most_common(14):
    print(f"  {v / tot * 100:6.2f} %  {v / 1e3:8.2f} ms  {cnt[n]:5d} x {v / cnt[n]:8.1f} us  {n}")
