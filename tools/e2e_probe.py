"""Where does the end-to-end step time go?  staging alone / step alone / overlapped pipeline."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
import routeformer_b200 as R
from routeformer_b200.parallel import DataParallelTrainer, BatchPrefetcher
dev = torch.device("cuda", 0)
cfg, spec, host_batch, host_targets = bench.build_case(64, 100, "gaze")
torch.manual_seed(0)
model = bench.build_model(cfg, spec, "gaze").to(dev).train()
lossf = R.FutureDiscountedLoss({0: 0.97}, epsilon=1.0, loss_function="smooth_l1")
trainer = DataParallelTrainer(model, lambda o, t: lossf(o[0], t[0]) + 0.5 * lossf(o[1], t[1]), use_cuda_graph=True)
pinned = {k: v.contiguous().pin_memory() for k, v in host_batch.items()}
pinned_t = tuple(t.contiguous().pin_memory() for t in host_targets)
batch = model.stage_batch(pinned, dev); targets = tuple(t.to(dev) for t in host_targets)
for _ in range(3): trainer.step(batch, targets)
torch.cuda.synchronize()
def timed(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("step only            %.2f ms" % timed(lambda: trainer.step(batch, targets).item()))
t0 = time.perf_counter(); model.stage_batch(pinned, dev, out=batch); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("stage: host enqueue %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print("stage only           %.2f ms" % timed(lambda: model.stage_batch(pinned, dev, out=batch)))
pf = BatchPrefetcher(model, dev, 2)
pf.submit(pinned, pinned_t)
def pipe():
    b, t = pf.get(); l = trainer.step(b, t); pf.release(b); pf.submit(pinned, pinned_t); return l.item()
pipe(); pipe()
print("pipeline             %.2f ms" % timed(pipe))
side = torch.cuda.Stream()
def manual():
    with torch.cuda.stream(side):
        model.stage_batch(pinned, dev, out=pf.slots[0][0])
    l = trainer.step(batch, targets); return l.item()
print("stage on side stream + independent step  %.2f ms" % timed(manual))
