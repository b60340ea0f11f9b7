"""The C++ host side above the C ABI (freeimpala_b200/host): the reference's Learner / SharedBuffer /
ModelManager interfaces and the threaded harness with the reference's flags (BASELINE.json configs[2])."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def _run(fi, *args, timeout=300):
    from freeimpala_b200 import build
    exe = build.HOST_BIN
    assert os.path.exists(exe), "host harness not built (python -m freeimpala_b200.build)"
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return r.returncode, (json.loads(line[-1]) if line else None), r.stderr


def test_threaded_harness_config3_shape(fi):
    """players 2, buffer-capacity 32, batch 32, 64 actor threads (configs[2]); T chosen so that every written
    trajectory is consumed: 64 agents x 4 iterations / 32 = 8 learner updates per player."""
    rc, out, err = _run(fi, "-p", 2, "-B", 32, "-M", 32, "-a", 64, "-T", 4, "-S", 100, "--game-steps", 100, "--agent-time", 0)
    assert rc == 0, err[-2000:]
    assert out["learner_updates"] == out["expected_updates"] == 16
    assert out["min_model_version"] == 1 + 8          # Model ctor -> 1, +1 per update (learner.h:40-45)
    assert out["kernel_launches"] > 0 and out["transitions_per_s"] > 0


def test_harness_checkpoints_and_resume(fi, tmp_path):
    d = str(tmp_path)
    rc, out, err = _run(fi, "-p", 1, "-B", 8, "-M", 4, "-a", 4, "-T", 4, "-S", 10, "--game-steps", 10, "--agent-time", 0,
                        "-c", 2, "-l", d)
    assert rc == 0, err[-2000:]
    names = sorted(os.listdir(d))
    assert "model_0_latest.bin" in names and "model_0_2.bin" in names and "model_0_4.bin" in names   # every 2 updates + final
    rc, out2, err = _run(fi, "-p", 1, "-B", 8, "-M", 4, "-a", 4, "-T", 2, "-S", 10, "--game-steps", 10, "--agent-time", 0, "-m", d)
    assert rc == 0, err[-2000:]
    assert out2["min_model_version"] == out["min_model_version"] + 2      # resumed from version 5, two more updates


def test_harness_rejects_invalid_parameters(fi):
    rc, out, err = _run(fi, "-M", 11, "-B", 10)                          # validateParameters: M <= B
    assert rc == 2 and "Batch size" in err
    rc, out, err = _run(fi, "--game-steps", 200, "-S", 100)
    assert rc == 2 and "Game steps" in err


def test_harness_batched_actor_inference(fi):
    rc, out, err = _run(fi, "-p", 1, "-B", 8, "-M", 4, "-a", 8, "-T", 2, "-S", 10, "--game-steps", 10, "--agent-time", 0,
                        "--infer-every", 1)
    assert rc == 0, err[-2000:]
    assert out["actor_inference_rows"] == 8 * 2 * 10


def test_reference_agent_runs_unmodified_against_the_shim(fi):
    """Drop-in proof: oracle/_ref/agent_dropin is the reference's UNMODIFIED include/freeimpala/agent.h (+ its
    metrics_tracker.h and the Buffer / BufferEntry / MessageTag / ELEMENT_SIZE definitions of its data_structures.h)
    compiled against freeimpala_b200/host/fi_host.hpp through the alias headers INTEGRATION.md describes
    (oracle/ref_shim/dropin/, built by `make -C oracle dropin` where /root/reference exists). 2 players x 2 reference
    Agents write rand() trajectories through SharedBuffer::write and pull weights through ModelManager while the learner's
    worker threads step on the GPU; the reference's own MetricsTracker counts the updates through the two calls
    fi_host::Learner::trainModel keeps (learner.h:34,48)."""
    from oracle import pyoracle as po
    if not os.path.exists(po.DROPIN_BIN):
        pytest.skip("oracle/_ref/agent_dropin was not built (needs /root/reference at build time)")
    r = subprocess.run([po.DROPIN_BIN, "2", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
    assert "learner model updates 8" in r.stdout      # 2 players x (2 agents x 8 iterations / batch 4) updates
