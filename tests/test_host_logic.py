"""Host logic of fi_host.hpp (the C++ mirror of the reference's SharedBuffer / ModelManager / Learner) on the CPU.

The header is compiled against a TEST DOUBLE of the C ABI (tests/host_double/fi_double.cpp: a host-memory FIFO and a step
that only bumps the version) -- the product library is not involved and has no CPU path. What is checked is the thread
logic the reference defines in include/freeimpala/learner.h:52-97 (checkpointModel, workerThread) and :158-197
(start / stop): iteration counting, drain on stop, checkpoint files, and that a failing step or readBatch stops the worker
instead of spinning. The same header runs against the real library on the GPU in tests/test_gpu_host.py."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("host_logic")
    exe = tmp / "host_logic"
    src = [os.path.join(HERE, "host_double", f) for f in ("host_logic_main.cpp", "fi_double.cpp")]
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wall", "-o", str(exe), *src], check=True)
    for d in ("run", "api"):
        (tmp / d).mkdir()
    out = subprocess.run([str(exe), str(tmp)], check=True, capture_output=True, text=True, timeout=120)
    return json.loads(out.stdout)


def test_worker_threads_count_iterations_and_publish_versions(report):
    r = report["run"]
    assert r["iterations"] == [r["expected"], r["expected"]]       # learner.h:72-97, one worker per player
    assert r["versions"] == [1 + r["expected"]] * 2                  # version + 1 per step (learner.h:40-45)
    assert r["consumed"] == 2 * 4 * r["expected"] and r["fifo"]      # M slots per iteration, FIFO per writer
    assert r["updates_counted"] == 2 * r["expected"]                 # recordLearnerModelUpdate point (learner.h:48)
    assert r["model_version"] == r["versions"][0] and r["model_bytes"] == 4096


def test_checkpoints_periodic_and_final_in_the_reference_format(report):
    r = report["run"]
    assert r["periodic_checkpoint"]                                   # every c iterations (learner.h:91-93)
    assert r["latest_file_version"] == r["versions"][0]               # final save in stop() (learner.h:187)
    assert r["latest_file_bytes"] == 8 + 4096                         # u64 version + raw bytes (data_structures.h:105-110)


def test_stop_drains_blocked_workers(report):
    d = report["drain"]
    assert d["stop_seconds"] < 2.0 and d["iterations"] == 0 and d["left_in_ring"] == 1


def test_failing_step_stops_the_worker(report):
    s = report["step_failure"]
    assert s["iterations"] == 2 and s["updates_counted"] == 2        # the failed step is not counted
    assert s["read_calls_while_idle"] == 0 and s["failed"]           # and the worker has left its loop; Learner::failed() says why


def test_failing_read_stops_the_worker_instead_of_spinning(report):
    s = report["read_failure"]
    assert s["iterations"] == 1 and s["read_calls"] == 2


def test_shared_buffer_and_model_manager_semantics(report):
    a = report["api"]
    assert a["oversize_write"] is False                               # data_structures.h:226,240
    assert a["try_writes"] == [True, True, False]                     # full ring -> false (:245-249)
    assert a["wait_without_update"] is False                          # waitForModelUpdate times out (:454-472)
    assert a["resumed_version"] == 2 and a["out_of_range_model"] is False


def test_host_logic_is_clean_under_thread_sanitizer(tmp_path):
    """The same scenarios built with -fsanitize=thread: no data race in fi_host.hpp's worker / checkpoint / stop logic
    (the reference relies on std::mutex + condition variables throughout)."""
    exe = tmp_path / "host_logic_tsan"
    src = [os.path.join(HERE, "host_double", f) for f in ("host_logic_main.cpp", "fi_double.cpp")]
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-o", str(exe), *src],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("libtsan is not available to this g++: " + build.stderr[-200:])
    for d in ("run", "api"):
        (tmp_path / d).mkdir()
    out = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in out.stderr:   # e.g. an unsupported address-space layout in the container
        pytest.skip(out.stderr[-200:])
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stderr[-2000:]
    assert json.loads(out.stdout)["run"]["fifo"]


REF = "/root/reference"
SPDLOG = "/opt/prime-rl/.venv/lib/python3.12/site-packages/flashinfer/data/spdlog/include"


@pytest.mark.parametrize("tsan", [False, True])
def test_reference_agent_runs_against_the_shim_on_the_cpu(tmp_path, tsan):
    """The reference's UNMODIFIED include/freeimpala/agent.h (2 players x 3 Agent threads, its own MetricsTracker) driven
    against fi_host.hpp exactly as in tests/test_gpu_host.py, but linked with the test double instead of the product
    library: the plumbing the reference's actors rely on (SharedBuffer::write from transfer threads created per iteration,
    ModelManager polling, agent.h:78-165) is exercised without a GPU, once more under ThreadSanitizer. Needs the reference
    tree (present in the build container, absent on the GPU box)."""
    root = os.path.dirname(HERE)
    if not os.path.exists(os.path.join(REF, "include", "freeimpala", "agent.h")) or not os.path.isdir(SPDLOG):
        pytest.skip("the reference tree / spdlog headers are not present here")
    exe = tmp_path / "agent_dropin_double"
    shim = os.path.join(root, "oracle", "ref_shim")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-w", "-o", str(exe), os.path.join(shim, "dropin_main.cpp"),
           os.path.join(HERE, "host_double", "fi_double.cpp"),
           f'-DFI_REF_DATA_STRUCTURES_H="{REF}/include/freeimpala/data_structures.h"', f'-DFI_REF_AGENT_H="{REF}/include/freeimpala/agent.h"',
           "-I" + os.path.join(shim, "dropin"), "-I" + os.path.join(REF, "include"), "-I" + SPDLOG,
           "-I" + os.path.join(root, "freeimpala_b200", "host"), "-I" + os.path.join(root, "include")]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and tsan:
        pytest.skip("libtsan is not available to this g++: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-3000:]
    out = subprocess.run([str(exe), "2", "3"], capture_output=True, text=True, timeout=300)
    if tsan and "FATAL: ThreadSanitizer" in out.stderr:
        pytest.skip(out.stderr[-200:])
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
    assert "learner model updates 12" in out.stdout      # 2 players x (3 agents x 8 iterations / batch 4) updates
    if tsan:
        ours = [b for b in out.stderr.split("WARNING: ThreadSanitizer")[1:] if "fi_host.hpp" in b or "fi_double.cpp" in b]
        assert not ours, ours[0][:3000]
