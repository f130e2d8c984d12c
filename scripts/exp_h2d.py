import torch, time, numpy as np
torch.cuda.set_device(0)
N = 4 << 30
def bw(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return N * reps / (time.perf_counter() - t0) / 1e9
dst = torch.empty(N, dtype=torch.uint8, device="cuda")
src_pin = torch.empty(N, dtype=torch.uint8, pin_memory=True); src_pin.fill_(1)
print("torch pin_memory, 1 stream      :", round(bw(lambda: dst.copy_(src_pin, non_blocking=True)), 1), "GB/s")
a = np.ones(N, dtype=np.uint8)
torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
ta = torch.from_numpy(a)
print("cudaHostRegister(numpy), 1 strm :", round(bw(lambda: dst.copy_(ta, non_blocking=True)), 1), "GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    h = N // 2
    with torch.cuda.stream(s1): dst[:h].copy_(src_pin[:h], non_blocking=True)
    with torch.cuda.stream(s2): dst[h:].copy_(src_pin[h:], non_blocking=True)
print("pin_memory, 2 streams           :", round(bw(two), 1), "GB/s")
def chunks():
    c = 64 << 20
    for o in range(0, N, c): dst[o:o+c].copy_(src_pin[o:o+c], non_blocking=True)
print("pin_memory, 64 MiB chunks       :", round(bw(chunks), 1), "GB/s")
back = torch.empty(N, dtype=torch.uint8, pin_memory=True)
print("D2H pin_memory                  :", round(bw(lambda: back.copy_(dst, non_blocking=True)), 1), "GB/s")
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -5; numactl -H 2>/dev/null | head -4; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
