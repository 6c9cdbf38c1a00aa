"""Runs the C++ Provider-mirror tests (k8s-runpod-kubelet_b200/host/host_test.cc): scenarios shaped like the
reference's annotations_test.go plus the batched tick bodies."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "k8s-runpod-kubelet_b200", "host")
BIN = os.path.join(HOST, "host_test")


def run(arg):
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", HOST, "-s"])
    out = subprocess.run([BIN, arg], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1].startswith("ok:")
    return out.stdout


def test_host_logic_cpu():
    """column producers, ports check, status translation table, record encoding -- no GPU involved"""
    import torch

    out = run("--cpu" if torch.cuda.is_available() else "--cpu-no-device")
    assert "0 failed" in out


@pytest.mark.gpu
def test_provider_over_cuda_engine():
    """CreatePod / ProcessPendingPods (one rpk_select per tick) / UpdateAllPodStatuses (one rpk_status_diff
    per tick) / NotifyPods / GetPod / GetPodStatus / DeletePod against a scripted RunPod API"""
    out = run("--gpu")
    assert "(gpu)" in out and "0 failed" in out
