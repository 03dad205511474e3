"""Developer tool (torchrun, one rank per GPU): bandwidth of peer-to-peer copies through symmetric memory -- pull (read the
peer's buffer) vs push (write into it), one copy vs the same bytes split over 2 / 4 streams -- plus the latency of the
symmetric-memory barrier and, for scale, an NCCL all-reduce of the same bytes."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nmax = 256 * 2 ** 20 // 4
    buf = symm.empty(nmax, dtype=torch.float32, device=dev)
    h = symm.rendezvous(buf, dist.group.WORLD)
    buf.fill_(rank)
    local_buf = torch.empty(nmax, dtype=torch.float32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    peer = (rank + 1) % world
    dist.barrier()
    for mb in (16, 64, 256):
        n = mb * 2 ** 20 // 4
        remote = h.get_buffer(peer, (n,), torch.float32, 0)
        for k in (1, 2, 4):
            def run(pull):
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                m = n // k
                for j in range(k):
                    s = streams[j]
                    s.wait_event(ev)
                    with torch.cuda.stream(s):
                        if pull:
                            local_buf[j * m:(j + 1) * m].copy_(remote[j * m:(j + 1) * m], non_blocking=True)
                        else:
                            remote[j * m:(j + 1) * m].copy_(local_buf[j * m:(j + 1) * m], non_blocking=True)
                    cur.wait_stream(s)
            t_pull = timed(lambda: run(True))
            t_push = timed(lambda: run(False))
            if rank == 0:
                print(f"{world} GPUs {mb:4d} MB over {k} stream(s): pull {mb / 1024 / t_pull * 1e3:7.1f} GB/s ({t_pull * 1e3:6.0f} us)  "
                      f"push {mb / 1024 / t_push * 1e3:7.1f} GB/s ({t_push * 1e3:6.0f} us)", flush=True)
        t_nccl = timed(lambda: dist.all_reduce(local_buf[:n]))
        if rank == 0:
            print(f"{world} GPUs {mb:4d} MB NCCL all-reduce {t_nccl * 1e3:6.0f} us (algorithmic {mb / 1024 / t_nccl * 1e3:6.1f} GB/s)", flush=True)
    t_bar = timed(lambda: h.barrier(0, 10000), 50)
    if rank == 0:
        print(f"symmetric-memory barrier {t_bar * 1e3:.1f} us", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
