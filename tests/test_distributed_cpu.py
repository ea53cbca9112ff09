"""Multi-process plumbing on CPU (gloo): TP, PP, PP x TP and EP layouts must reproduce the single-process
tokens of the same model (same global weights sharded through the loader)."""
from conftest import scratch_dir
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(pp, tp, arch="Qwen3ForCausalLM", method="chunked_prefill", port=29811):
    out = os.path.join(scratch_dir("gllm_b200_f_"), "f" + ".json")
    env = dict(os.environ, PYTHONPATH=ROOT, GLLM_B200_LOG="WARNING")
    script = os.path.join(ROOT, "tests", "mp_engine_cpu.py")
    n = pp * tp
    if n == 1:
        cmd = [sys.executable, script, "1", "1", out, arch, method]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), script, str(pp), str(tp), out, arch, method]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and os.path.exists(out), r.stdout[-2000:] + r.stderr[-3000:]
    _run.last_out = out
    with open(out) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def single():
    return _run(1, 1)


def test_tp2_matches_single(single):
    assert _run(1, 2, port=29821) == single


def test_pp2_matches_single(single):
    assert _run(2, 1, port=29831) == single


def test_pp2_tp2_token_throttling_matches_single(single):
    assert _run(2, 2, method="token_throttling", port=29841) == single


def test_mixtral_ep2_matches_single():
    ref = _run(1, 1, arch="MixtralForCausalLM")
    assert _run(1, 2, arch="MixtralForCausalLM", port=29851) == ref


def test_pp2_tile_streamed_recv_matches_single(single, monkeypatch):
    """Force tiny row tiles so the stage-1 input really arrives as several NCCL/gloo p2p tiles and the
    first layer's add+norm / QKV projection run per tile."""
    monkeypatch.setenv("GLLM_PP_TILE_ROWS", "8")
    assert _run(2, 1, port=29871) == single


def test_tp2_async_lookahead_matches_single(single, monkeypatch):
    """Async scheduling under TP: the lookahead batch (placeholder tokens + device-side feed indices) travels to
    the peer rank over the packed ZeroMQ frame and every rank feeds from its own sampler output."""
    monkeypatch.setenv("GLLM_TEST_ASYNC", "1")
    assert _run(1, 2, port=29881) == single


def test_pp2_tp2_sharded_stage_transfer_matches_single(single, monkeypatch):
    """PP x TP: each TP rank ships only its 1/tp slice of a stage-boundary tile and the next stage all-gathers
    it (uneven slices, several tiles per micro-batch, decode steps stay replicated)."""
    monkeypatch.setenv("GLLM_PP_TILE_ROWS", "7")
    monkeypatch.setenv("GLLM_PP_SHARD_MIN_ROWS", "3")
    monkeypatch.setenv("GLLM_TEST_PP_STATS", "1")
    assert _run(2, 2, port=29891) == single
    stats = []
    for rank in (2, 3):                                   # the second stage's TP ranks
        with open(f"{_run.last_out}.pp{rank}") as f:
            stats.append(json.load(f))
    assert all(st["sharded_tiles"] > 0 and st["replicated_tiles"] > 0 for st in stats), stats


@pytest.mark.parametrize("which", ["qwen2_5", "qwen3"])
def test_vision_tower_tp2_matches_replicated(which):
    """Heads / MLP-intermediate sharded vision towers == the replicated tower, on both ranks."""
    out = os.path.join(scratch_dir("gllm_b200_f_"), "vision.json")
    env = dict(os.environ, PYTHONPATH=ROOT, GLLM_B200_LOG="WARNING")
    port = 29901 if which == "qwen2_5" else 29911
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_vision_tp_cpu.py"), out, which]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and os.path.exists(out), r.stdout[-2000:] + r.stderr[-3000:]
    with open(out) as f:
        res = json.load(f)
    assert res["ok"], res


@pytest.mark.parametrize("pp,tp,port", [(1, 2, 29921), (2, 2, 29931)])
def test_shared_memory_batch_ring_matches_single(single, monkeypatch, pp, tp, port):
    """GLLM_BATCH_TRANSPORT=shm: the driver's batches (and control messages) reach the other ranks through the
    shared-memory broadcast ring instead of ZeroMQ sockets."""
    monkeypatch.setenv("GLLM_BATCH_TRANSPORT", "shm")
    assert _run(pp, tp, port=port) == single


def test_tp2_sharding_edge_shapes_match_single(monkeypatch):
    """Multi-query attention (1 KV head replicated on both ranks), a vocabulary that is not a multiple of the
    shard padding, tied embeddings, an odd intermediate size: TP2 == single process."""
    monkeypatch.setenv("GLLM_TEST_CFG", json.dumps({"num_key_value_heads": 1, "vocab_size": 777,
                                                    "tie_word_embeddings": True, "intermediate_size": 264}))
    ref_tokens = _run(1, 1)
    assert _run(1, 2, port=29941) == ref_tokens
    assert _run(2, 2, port=29951) == ref_tokens      # tied embeddings live on the first AND the last stage


def test_pp2_explicit_layer_assignment_matches_single(single, monkeypatch):
    """`--assigned-layers 3,1` (reference: dist_utils.py:175-209): an uneven explicit split of the 4 layers."""
    monkeypatch.setenv("GLLM_TEST_ASSIGNED", "3,1")
    assert _run(2, 1, port=29971) == single


def test_mixtral_tp2_without_expert_parallelism_matches_single(monkeypatch):
    """`--disable-ep`: every rank holds all experts with intermediate / tp columns (reference: layer.py:249-259)."""
    ref_tokens = _run(1, 1, arch="MixtralForCausalLM")
    monkeypatch.setenv("GLLM_TEST_NO_EP", "1")
    assert _run(1, 2, arch="MixtralForCausalLM", port=29991) == ref_tokens


def test_tp2_vocab_parallel_sampling_runs_and_keeps_greedy_rows_exact(monkeypatch):
    """SURVEY §2.4 X4 under gloo: a batch mixing greedy, top-k/top-p + repetition-penalty and unfiltered temperature
    rows never gathers the [E, V] logits at TP2 (the runner counts vocab-parallel sampling steps); its greedy rows
    — one of them with a repetition penalty — reproduce the single-process tokens."""
    monkeypatch.setenv("GLLM_TEST_SAMPLED", "1")
    ref = _run(1, 1)
    assert _run(1, 2, port=29901) == ref


def test_second_engine_in_the_same_processes_still_reaches_its_peers(single, monkeypatch):
    """Round-2 hardware bug: engines of one torchrun job re-used the same ipc file names; ZeroMQ closes listeners
    asynchronously and unlinks the file when it finally does, which could be after the next engine had bound the path —
    its peers then never heard from the driver. Endpoints are per engine instance now."""
    monkeypatch.setenv("GLLM_TEST_TWO_ENGINES", "1")
    assert _run(1, 2, port=29911) == single
    assert _run(2, 1, port=29921) == single
