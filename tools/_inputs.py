"""Inputs of a bench workload for the one-off tools: (task(fm, packed, n_raw), packed host records, workload dict)."""
import sys
sys.path.insert(0, ".")
import torch
import bench


def workload(name="C2-small", M=None, P=None):
    w = dict(bench.WORKLOADS[name])
    if M:
        w["M"] = M
    if P:
        w["P"] = P
    dev = torch.device("cuda:0")
    packed = bench.make_packed(w["n"], w["M"], 7, dev)
    task = (bench.task_chol if w["gen"] == "chol" else bench.task_structured)(w, dev)[0]
    return task, packed, w
