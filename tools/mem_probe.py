import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200quant
from b200quant.harness import ResNetInt8
dev = torch.device("cuda", 0)
for B in (32, 128, 256):
    torch.cuda.reset_peak_memory_stats()
    m = ResNetInt8().to(dev)
    opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
    x = torch.randn(B, 3, 224, 224, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
    for i in range(3):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(m(x), y); loss.backward(); opt.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(m(x), y); loss.backward(); opt.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(B, "peak GB", torch.cuda.max_memory_allocated() / 1e9, "alloc after GB", torch.cuda.memory_allocated() / 1e9, "ms", dt * 1e3, "img/s", B / dt, flush=True)
    del m, opt, x, y, loss
    torch.cuda.empty_cache()
    print(" after free GB", torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9, flush=True)
