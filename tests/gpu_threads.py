"""Threading contract of the C ABI (nn.DataParallel's parallel_apply, train.py:493): the same library is
called concurrently from several host threads, one per device, each on its own stream.  Diagnostic for a multi-GPU box:
    python tests/gpu_threads.py        (uses every visible GPU; with one GPU it runs two threads on two streams)"""
import os, sys, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosnet_b200 import coattention_forward_raw as op
from oracle import coattn_oracle as orc

ndev = torch.cuda.device_count()
workers = max(2, ndev)
n, h, w = 4, 40, 40
inputs, want = [], []
for k in range(workers):
    v_a, v_b = (torch.from_numpy(x) for x in orc.synthetic_features(100 + k, n, h, w, 0.66))
    W, g, b = (torch.from_numpy(x) for x in orc.synthetic_weights(200 + k, bias=True))
    inputs.append((v_a, v_b, W, g, b))
# reference results: one device, one thread
dev0 = torch.device("cuda:0")
for args in inputs:
    out = op(*(t.to(dev0) for t in args), want_z=False)
    want.append([t.cpu() for t in out[:2]])
torch.cuda.synchronize()
errors = []

def work(k):
    try:
        dev = torch.device("cuda", k % ndev)
        torch.cuda.set_device(dev)
        stream = torch.cuda.Stream(dev)
        args = [t.to(dev) for t in inputs[k]]
        with torch.cuda.stream(stream):
            for _ in range(20):
                out = op(*args, want_z=False)
        stream.synchronize()
        for got, ref in zip(out[:2], want[k]):
            if not torch.equal(got.cpu(), ref):
                errors.append(f"worker {k} on {dev}: result differs from the single-threaded run")
    except Exception as e:  # noqa: BLE001
        errors.append(f"worker {k}: {e!r}")

threads = [threading.Thread(target=work, args=(k,)) for k in range(workers)]
for t in threads: t.start()
for t in threads: t.join()
print(f"{workers} threads on {ndev} device(s):", "OK (bit-identical to the single-threaded results)" if not errors else errors)
sys.exit(1 if errors else 0)
