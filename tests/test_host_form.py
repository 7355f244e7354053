"""Host-side choice of the batch-fit form (fpb_engine.cu, fit_slice_host): no GPU needed.

Eight ranks copying through one host make the host the bound (23 GB/s per GPU instead of 55): the wavefront form loses
to the chunk-per-engine form there (298.8 vs 267.7 ms per step, profiles/r02_e2e.md), so the library counts the ranks the
launcher put on the node and switches above 2 (4 GPUs: 188.8 vs 182.9 ms)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = ("import fitpack_b200._lib as L; l = L.lib(); "
        "print(l.fpb_host_sharers(), l.fpb_host_form(1).decode(), l.fpb_host_form(3000).decode(), "
        "l.fpb_host_form(100000).decode(), l.fpb_host_form(1 << 20).decode())")


def _run(env_extra):
    env = {k: v for k, v in os.environ.items()
           if k not in ("LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE", "FPB_LOCAL_RANKS",
                        "FPB_HOST_MODE", "FPB_SOLO")}
    env.update(env_extra)
    out = subprocess.check_output([sys.executable, "-c", CODE], cwd=ROOT, env=env, text=True)
    return out.split()


def test_one_rank_takes_the_wavefront_form_for_large_batches():
    assert _run({}) == ["1", "warp-per-curve", "warp-per-curve", "chunks", "wavefront"]


def test_eight_ranks_per_node_take_the_chunk_per_engine_form():
    for var in ("LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE", "FPB_LOCAL_RANKS"):
        assert _run({var: "8"}) == ["8", "warp-per-curve", "warp-per-curve", "chunks", "chunks"], var
    assert _run({"LOCAL_WORLD_SIZE": "4"})[-1] == "chunks"
    assert _run({"LOCAL_WORLD_SIZE": "2"})[-1] == "wavefront"
    assert _run({"LOCAL_WORLD_SIZE": "8", "FPB_LOCAL_RANKS": "2"})[0] == "2"        # the explicit knob wins


def test_overrides():
    assert _run({"LOCAL_WORLD_SIZE": "8", "FPB_HOST_MODE": "wave"})[-1] == "wavefront"
    assert _run({"FPB_HOST_MODE": "workers"})[-1] == "chunks"
    assert _run({"FPB_SOLO": "0"})[1:3] == ["chunks", "chunks"]
