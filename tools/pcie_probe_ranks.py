"""Host<->device copy ceiling of a multi-GPU box, all ranks copying at the same time (context for bench.py's e2e at N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/pcie_probe_ranks.py

Every rank pins 2 x 256 MiB of host memory and times H2D, D2H and both directions at once between two barriers (CUDA events,
max over ranks).  Done twice: with the process where the launcher put it, and with the process (and therefore its pinned
pages: first touch) moved to the CPUs of its GPU's NUMA node.  Rank 0 prints one JSON line per mode."""
import json
import os

import torch
import torch.distributed as dist


def numa_node_of_gpu(index):
    """(NUMA node or -1, PCI address) of a GPU; virtual machines often expose neither a node nor more than one node."""
    try:
        props = torch.cuda.get_device_properties(index)
        bus = f"{int(props.pci_domain_id):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read()), bus
    except Exception as e:  # noqa: BLE001 - a probe must not die on a missing sysfs entry
        return -1, f"unknown ({type(e).__name__})"


def cpus_of_node(node):
    try:
        text = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except OSError:
        return set()
    cpus = set()
    for part in text.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def measure(device, world):
    n = 256 << 20
    host_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host_a.fill_(1)
    host_b.fill_(2)
    dev_a = torch.empty(n, dtype=torch.uint8, device=device)
    dev_b = torch.empty(n, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            dev_a.copy_(host_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            host_b.copy_(dev_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {}
    for name, fn, factor in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both_sum", both, 2)):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 12
        for _ in range(reps):
            fn()
        s1.synchronize()
        s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps / 1e3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_gbs_per_gpu"] = factor * n / float(t.item()) / 1e9
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    node, bus = numa_node_of_gpu(local)
    before = sorted(os.sched_getaffinity(0))
    info = {"rank": rank, "gpu_bus": bus, "gpu_numa_node": node, "cpus_allowed": len(before)}
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, info)
    else:
        gathered = [info]
    first = measure(device, world)
    local_cpus = cpus_of_node(node) & set(before) if node >= 0 else set()
    moved = False
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)
        moved = True
    second = measure(device, world)
    if rank == 0:
        try:
            quota = open("/sys/fs/cgroup/cpu.max").read().strip()
        except OSError:
            quota = "?"
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
        print(json.dumps({"probe": "pcie_ranks", "world": world, "ranks": gathered, "numa_nodes": nodes, "cgroup_cpu_max": quota,
                          "os_cpu_count": os.cpu_count()}))
        print(json.dumps({"mode": "as launched", **{k: round(v, 2) for k, v in first.items()},
                          "aggregate_both_sum_gbs": round(first["both_sum_gbs_per_gpu"] * world, 1)}))
        print(json.dumps({"mode": "process and pinned pages on the GPU's NUMA node" if moved else "NUMA node unknown: unchanged",
                          **{k: round(v, 2) for k, v in second.items()},
                          "aggregate_both_sum_gbs": round(second["both_sum_gbs_per_gpu"] * world, 1)}))
    if world > 1:
        dist.destroy_process_group()


main()
