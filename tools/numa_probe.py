#!/usr/bin/env python
"""Developer tool: host-side topology of a GPU box and the D2H bandwidth into pinned memory placed on each NUMA node.
usage: python tools/numa_probe.py [device]"""
import ctypes
import os
import subprocess
import sys
import time

import torch

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'; cat /sys/devices/system/node/online; "
                     "nproc; cat /proc/self/status | grep -i 'cpus_allowed_list\\|mems_allowed_list'",
                     shell=True, capture_output=True, text=True).stdout)
bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
print(q)
for line in q.strip().splitlines():
    idx, busid = [s.strip() for s in line.split(",")]
    p = f"/sys/bus/pci/devices/{busid[4:].lower()}/numa_node"
    print(idx, busid, "numa_node", open(p).read().strip() if os.path.exists(p) else "?")
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238   # x86-64
MPOL_DEFAULT, MPOL_BIND = 0, 2
nodes = [int(n[4:]) for n in os.listdir("/sys/devices/system/node") if n.startswith("node") and n[4:].isdigit()]
torch.cuda.set_device(dev)
src = torch.empty(400 << 20, dtype=torch.uint8, device="cuda")
for node in sorted(nodes) + [None]:
    if node is None:
        rc = libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), 64)
    if rc != 0:
        print("node", node, "set_mempolicy failed", os.strerror(ctypes.get_errno()))
        continue
    try:
        dst = torch.empty(400 << 20, dtype=torch.uint8, pin_memory=True)
    except Exception as e:  # noqa: BLE001
        print("node", node, "alloc failed", e)
        continue
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"pinned on node {node}: D2H {400 * 1.048576 / 1e3 / dt:.1f} GB/s")
    del dst
libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
