"""2-rank NCCL ping-pong latency of a 96-byte message (the carrier-phase hand-off), idle and beside a busy kernel."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
x = torch.zeros(12, dtype=torch.float64, device="cuda")
def hops(n):
    for i in range(n):
        if rank == 0:
            dist.send(x, 1); dist.recv(x, 1)
        else:
            dist.recv(x, 0); dist.send(x, 0)
hops(20); torch.cuda.synchronize()
t = time.perf_counter(); hops(200); torch.cuda.synchronize()
if rank == 0: print("idle: %.1f us per hop" % ((time.perf_counter() - t) / 400 * 1e6))
side = torch.cuda.Stream(); a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
with torch.cuda.stream(side):
    for _ in range(200): b = a @ a
t = time.perf_counter(); hops(200); torch.cuda.synchronize(torch.cuda.current_stream())
torch.cuda.current_stream().synchronize()
if rank == 0: print("beside matmuls: %.1f us per hop" % ((time.perf_counter() - t) / 400 * 1e6))
torch.cuda.synchronize(); dist.destroy_process_group()
