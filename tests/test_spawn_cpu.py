"""Spawned-worker mode (front-end process + ZeroMQ + one worker process per rank) and the HTTP server as a
real subprocess driven by the serving benchmark client — all on CPU."""
from conftest import scratch_dir
import json
import os
import socket
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("async_worker", [False, True])
def test_spawn_mode_pp2_generates(async_worker):
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
if __name__ == "__main__":
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    llm = LLM(tiny("Qwen3ForCausalLM", num_hidden_layers=4), load_format="dummy", pp_size=2, tp_size=1, maxp=48,
              maxd=16, num_cpu_pages=128, model_max_length=256, log_stats=False, device="cpu",
              master_port={_free_port()}, schedule_method="token_throttling", use_async_worker={async_worker})
    outs = llm.generate(tokens=[[5, 17, 99], [9] * 40, list(range(20, 120))], output_lens=[6, 7, 8], ignore_eos=True)
    assert [len(s.token_ids) - s.prompt_len for s in outs] == [6, 7, 8]
    print("SPAWN_OK")
    llm.shutdown()
"""
    f = os.path.join(scratch_dir("gllm_b200_f_"), "f" + ".py")
    open(f, "w").write(code)
    r = subprocess.run([sys.executable, f], capture_output=True, text=True, timeout=240,
                       env=dict(os.environ, GLLM_B200_LOG="WARNING"))
    assert "SPAWN_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_server_subprocess_with_serving_benchmark():
    import requests
    port = _free_port()
    cfg = scratch_dir("gllm_b200_srv_")
    from gllm_b200.models.presets import tiny
    with open(os.path.join(cfg, "config.json"), "w") as f:
        json.dump(tiny("Qwen3ForCausalLM", max_position_embeddings=2048), f)
    env = dict(os.environ, PYTHONPATH=ROOT, GLLM_B200_LOG="WARNING")
    srv = subprocess.Popen([sys.executable, "-m", "gllm_b200.entrypoints.api_server", "--model-path", cfg,
                            "--load-format", "dummy", "--port", str(port), "--host", "127.0.0.1", "--maxp", "64",
                            "--maxd", "32", "--model-max-length", "1100", "--enable-prefix-caching"],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        for _ in range(120):
            try:
                if requests.get(f"http://127.0.0.1:{port}/health", timeout=1).status_code == 200:
                    break
            except Exception:  # noqa: BLE001
                time.sleep(0.5)
        else:
            srv.kill()
            raise AssertionError("server did not come up: " + srv.stdout.read()[-3000:])
        out = os.path.join(scratch_dir("gllm_b200_f_"), "f" + ".json")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "benchmark_serving.py"), "--port",
                            str(port), "--num-prompts", "12", "--request-rate", "50", "--vocab-size", "500",
                            "--max-output-len", "8", "--save-result", out, "--goodput", "ttft:60000",
                            "--arrival-stage", "3", "--stage-interval", "0.2"],     # staged arrivals (3 x 4 requests)
                           capture_output=True, text=True, timeout=240, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res = json.load(open(out))
        assert res["completed"] == 12 and res["failed"] == 0 and res["arrival_stage"] == 3
        for k in ("median_ttft_ms", "median_tpot_ms", "median_itl_ms", "median_e2el_ms", "output_throughput",
                  "request_goodput", "p99_ttft_ms"):
            assert k in res, k
        r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "benchmark_prefix_serving.py"), "--port",
                            str(port), "--num-users", "3", "--rounds", "3", "--vocab-size", "500", "--system-len",
                            "64", "--turn-len", "16", "--answer-len", "4"],
                           capture_output=True, text=True, timeout=240, env=env)
        assert r.returncode == 0 and "cache_hit_rate" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        hit = float(r.stdout.strip().splitlines()[-1].split()[-1])
        assert hit > 0.0
    finally:
        srv.terminate()
        try:
            srv.wait(timeout=10)
        except Exception:  # noqa: BLE001
            srv.kill()


def test_master_slave_two_node_launch_on_localhost():
    """`--launch-mode master|slave`: two "nodes" (one worker rank each, PP=2) talk over TCP ZeroMQ + gloo; every
    TCP endpoint is bound on the master node, the slave only needs the master address
    (reference: api_server.py:312-322, comm.py)."""
    mp_, zp = _free_port(), _free_port()
    common = f"""
import sys
sys.path.insert(0, {ROOT!r})
from gllm_b200 import LLM
from gllm_b200.models.presets import tiny
cfg = tiny("Qwen3ForCausalLM", num_hidden_layers=4)
kw = dict(load_format="dummy", pp_size=2, tp_size=1, maxp=48, maxd=16, num_cpu_pages=128, model_max_length=256,
          log_stats=False, device="cpu", master_addr="127.0.0.1", master_port={mp_}, zmq_port_base={zp},
          host="127.0.0.1")
"""
    slave = common + """
if __name__ == "__main__":
    llm = LLM(cfg, launch_mode="slave", worker_ranks=[1], **kw)
    for p in llm.procs:
        p.join()
"""
    master = common + """
if __name__ == "__main__":
    llm = LLM(cfg, launch_mode="master", worker_ranks=[0], **kw)
    outs = llm.generate(tokens=[[5, 17, 99], [9] * 40], output_lens=[6, 7], ignore_eos=True)
    assert [len(s.token_ids) - s.prompt_len for s in outs] == [6, 7]
    print("MULTINODE_OK")
    llm.shutdown()
"""
    fs, fm = os.path.join(scratch_dir("gllm_b200_f_"), "f" + "_s.py"), os.path.join(scratch_dir("gllm_b200_f_"), "f" + "_m.py")
    open(fs, "w").write(slave)
    open(fm, "w").write(master)
    env = dict(os.environ, GLLM_B200_LOG="WARNING")
    ps = subprocess.Popen([sys.executable, fs], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        time.sleep(1.0)
        pm = subprocess.run([sys.executable, fm], env=env, capture_output=True, text=True, timeout=240)
        assert "MULTINODE_OK" in pm.stdout, pm.stdout[-1500:] + pm.stderr[-2500:]
    finally:
        ps.terminate()
        try:
            ps.wait(timeout=10)
        except Exception:  # noqa: BLE001
            ps.kill()


def test_fault_injection_front_end_exits_non_zero():
    """Fail-stop contract: a worker that dies marks itself dead and the front-end process exits non-zero
    (reference: llm_engine.py:108-117, worker.py:239-249). `GLLM_FAULT_INJECT=<rank>:<step>` kills rank 1 of a
    PP=2 engine after a few steps."""
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
if __name__ == "__main__":
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    llm = LLM(tiny("Qwen3ForCausalLM", num_hidden_layers=4), load_format="dummy", pp_size=2, tp_size=1, maxp=48,
              maxd=16, num_cpu_pages=128, model_max_length=256, log_stats=False, device="cpu",
              master_port={_free_port()})
    llm.generate(tokens=[[5, 17, 99], [9] * 40], output_lens=[60, 60], ignore_eos=True)
    print("SHOULD_NOT_REACH")
"""
    f = os.path.join(scratch_dir("gllm_b200_f_"), "f" + ".py")
    open(f, "w").write(code)
    env = dict(os.environ, GLLM_B200_LOG="WARNING", GLLM_FAULT_INJECT="1:5")
    r = subprocess.run([sys.executable, f], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode != 0 and "SHOULD_NOT_REACH" not in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
    assert "fault injected" in (r.stdout + r.stderr)


def test_shutdown_removes_ipc_socket_files(tmp_path):
    """Engine shutdown must not leak ipc socket files (workers that get terminated cannot unlink their own)."""
    import glob
    from gllm_b200.engine.comm import Comm, ipc_base
    base = ipc_base()
    root = base[len("ipc://"):]
    fe = Comm(base, -1, 2, 0, frontend=True).init()
    drv = Comm(base, 0, 2, 0).init()
    peer = Comm(base, 1, 2, 0).init()
    assert len(glob.glob(root + "_*")) >= 3
    peer.close()
    assert not glob.glob(root + "_batch_1")
    drv.close()
    fe.close(unlink_all=True)
    assert glob.glob(root + "_*") == []


def test_offline_throughput_benchmark_and_batch_example():
    """benchmarks/benchmark_throughput.py and examples/batch_inference.py end to end on a tiny dummy model
    (reference: benchmarks/benchmark_throughput.py, examples/batch_inference.py)."""
    cfg = scratch_dir("gllm_b200_tp_")
    from gllm_b200.models.presets import tiny
    with open(os.path.join(cfg, "config.json"), "w") as f:
        json.dump(tiny("Qwen3ForCausalLM", max_position_embeddings=4096), f)
    env = dict(os.environ, PYTHONPATH=ROOT, GLLM_B200_LOG="WARNING")
    out = os.path.join(cfg, "throughput.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "benchmark_throughput.py"), "--model-path", cfg,
                        "--load-format", "dummy", "--num-prompts", "6", "--maxp", "64", "--maxd", "16",
                        "--output-json", out], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "Throughput:" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    assert res["num_prompts"] == 6 and res["output_tokens"] > 0 and res["output_tokens_per_s"] > 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "batch_inference.py"), "--model-path", cfg,
                        "--load-format", "dummy", "--num-prompts", "4", "--output-len", "6"],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "output tok/s" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_mmlu_pro_script_and_chat_clients_against_server():
    """benchmarks/evaluate_MMLU_pro.py (local json dataset) and the example chat clients against a real server
    process with a tokenizer + chat template (reference: benchmarks/evaluate_MMLU_pro.py, examples/client.py)."""
    import requests
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_api_cpu import _make_model_dir
    model = _make_model_dir()
    data = scratch_dir("gllm_b200_mmlu_")
    rows = [{"question": f"how are you w{i} ?", "options": ["hello", "world", "the", "a"], "answer": "ABCD"[i % 4],
             "category": "math" if i % 2 else "law", "cot_content": ""} for i in range(8)]
    with open(os.path.join(data, "test.json"), "w") as f:
        json.dump(rows, f)
    with open(os.path.join(data, "validation.json"), "w") as f:
        json.dump(rows[:4], f)
    port = _free_port()
    env = dict(os.environ, PYTHONPATH=ROOT, GLLM_B200_LOG="WARNING")
    srv = subprocess.Popen([sys.executable, "-m", "gllm_b200.entrypoints.api_server", "--model-path", model,
                            "--port", str(port), "--host", "127.0.0.1", "--maxp", "64", "--maxd", "16",
                            "--model-max-length", "250"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
    try:
        for _ in range(120):
            try:
                if requests.get(f"http://127.0.0.1:{port}/health", timeout=1).status_code == 200:
                    break
            except Exception:  # noqa: BLE001
                time.sleep(0.5)
        else:
            srv.kill()
            raise AssertionError("server did not come up: " + srv.stdout.read()[-3000:])
        r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "evaluate_MMLU_pro.py"), "--data-dir", data,
                            "--port", str(port), "--num-per-subject", "3", "--max-tokens", "4", "--workers", "2"],
                           capture_output=True, text=True, timeout=240, env=env)
        assert r.returncode == 0 and "overall" in r.stdout and "(6)" in r.stdout.splitlines()[-1], \
            r.stdout[-2000:] + r.stderr[-2000:]
        for extra in ([], ["--stream"]):     # examples/client.py: plain and SSE completions
            r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "client.py"), "--port", str(port),
                                "--prompt", "hello world how are you", "--max-tokens", "5", *extra],
                               capture_output=True, text=True, timeout=120, env=env)
            assert r.returncode == 0 and r.stdout.strip(), r.stdout[-1000:] + r.stderr[-2000:]
        # examples/chat_client.py is interactive: one question, then EOF
        r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "chat_client.py"), "--port", str(port)],
                           input="hello world\n", capture_output=True, text=True, timeout=120, env=env)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    finally:
        srv.terminate()
        try:
            srv.wait(timeout=10)
        except Exception:  # noqa: BLE001
            srv.kill()


def test_bench_reference_arm_contract_without_gpu():
    """`bench.py --impl reference` must always exit 0 with ONE JSON line: either the reference's numbers or
    {"impl": "reference", "unavailable": <why>} (here: no GPU, so the reference's CUDA extensions cannot load)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and ("unavailable" in d or "value" in d)
    # ranks other than 0 (torchrun launches N of them) stay silent: the reference spawns its own workers
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert r.returncode == 0 and not r.stdout.strip()
