"""CPU tests of the host-side logic: the C-ABI library exports what include/local_icp.h declares,
the generator, the sharding, and the world_size-2 gather over gloo."""
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    native = importlib.import_module("4d_plant_analysis_b200._native")
    native.build()
    lib = native.load()
    header = open(os.path.join(ROOT, "include", "local_icp.h")).read()
    declared = set(re.findall(r"\b(lip_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in local_icp.h but not exported"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert lib.lip_compiled_arch() == 100
    p = native.LipParams()
    lib.lip_default_params(__import__("ctypes").byref(p))
    assert (p.max_iter, p.cost_rise_tol, p.cost_flat_tol, p.obtuse_eps, p.cos_den_eps) == (500, 1e-4, 0.01, 1e-5, 1e-5)
    assert p.reject_dist == float("inf")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = importlib.import_module("4d_plant_analysis_b200")
    with pytest.raises(L.NativeError):
        L.default_context()
    x = np.random.default_rng(0).normal(size=(20, 3))
    with pytest.raises(L.NativeError):
        L.my_icp(x, x)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "4d_plant_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_generator_is_deterministic_and_shaped(syn):
    a, b = syn.make_config(1), syn.make_config(1)
    np.testing.assert_array_equal(a.xyz1, b.xyz1)
    np.testing.assert_array_equal(a.lm2, b.lm2)
    assert a.n_segments == 30 and 15000 < a.off1[-1] < 26000
    K = np.diff(a.lm_off)
    assert K.min() >= 2 and K.max() <= 20
    c3 = syn.make_config(3, scale=0.5)
    assert (np.diff(c3.lm_off) == 1).any()
    c4 = syn.make_config(4, scale=0.1)
    assert c4.n_frame_pairs == 15
    sub = a.take([3, 7])
    np.testing.assert_array_equal(sub.segment(1)[0], a.segment(7)[0])


def test_sharding():
    D = importlib.import_module("4d_plant_analysis_b200.distributed")
    costs = np.array([5.0, 1, 1, 1, 4, 3, 2, 2])
    sh = D.shard_segments(costs, 3)
    assert sorted(np.concatenate(sh).tolist()) == list(range(8))
    loads = [costs[s].sum() for s in sh]
    assert max(loads) - min(loads) <= 2
    fp = np.repeat(np.arange(15), 2)
    sh = D.shard_segments(np.ones(30), 8, fp)
    assert [len(s) // 2 for s in sh] == [2, 2, 2, 2, 2, 2, 2, 1]
    assert all(np.array_equal(x, np.arange(x[0], x[-1] + 1)) for x in sh)      # contiguous blocks: shared frames stay on a rank
    # the re-assembly indices of the packed gather address every entry exactly once
    off1, off2, lmo = np.array([0, 5, 9, 12]), np.array([0, 4, 10, 11]), np.array([0, 2, 3, 5])
    plan = D.make_shard_plan(off1, off2, lmo, 2)
    allidx = np.concatenate([plan.index21, plan.index12, plan.index_cnt])
    assert len(np.unique(allidx)) == len(allidx) and allidx.max() < 2 * plan.pad
    assert plan.sizes[:, 0].sum() == 11 and plan.sizes[:, 1].sum() == 12 and plan.sizes[:, 2].sum() == 3


def test_series_batch_views_and_subseries(syn):
    """a growth series stores every frame once; as_batch() gathers what the batched entry point would upload; a sub-series
    keeps exactly the meshes its frame pairs use"""
    ser = syn.make_series(1, n_frames=5, scale=0.1)
    b = ser.as_batch()
    assert ser.n_frame_pairs == 4 and ser.n_segments == 4 * 3 and b.n_correspondences == ser.n_correspondences
    assert len(ser.frames) < len(b.xyz1) + len(b.xyz2)                       # interior frames are shared
    # frame k is mesh2 of frame pair k-1 and mesh1 of frame pair k
    for s in range(3, 6):
        np.testing.assert_array_equal(b.segment(s)[0], b.segment(s - 3)[1])
    sub = ser.take_frame_pairs([1, 2])
    sb, ref = sub.as_batch(), b.take(np.nonzero(np.isin(b.frame_pair, [1, 2]))[0])
    np.testing.assert_array_equal(sb.xyz1, ref.xyz1); np.testing.assert_array_equal(sb.xyz2, ref.xyz2)
    np.testing.assert_array_equal(sb.lm1, ref.lm1); np.testing.assert_array_equal(sb.lm_off, ref.lm_off)
    assert len(sub.frames) == sum(len(b.segment(s)[0]) for s in range(3, 6)) + sum(len(b.segment(s)[1]) for s in range(3, 9))


_WORKER = r'''
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.environ["LIP_ROOT"])
import torch, torch.distributed as dist
dist.init_process_group("gloo")
D = importlib.import_module("4d_plant_analysis_b200.distributed")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
from oracle import local_icp_oracle as O      # the CPU checker stands in for the GPU compute in this host-logic test
def compute(sub):
    T21, T12, cnt = [], [], []
    for s in range(sub.n_segments):
        a, c = O.segment_pair_maps(*sub.segment(s), nn="brute")
        T21.append(a); T12.append(c)
        cnt.append([len(a), len(c), int(a.sum() % 1000), 7, int(c.sum() % 1000), 2 ** 40 + len(a)])      # stand-in int64 score counts
    return np.concatenate(T21), np.concatenate(T12), np.array(cnt, np.int64)
# (1) one frame pair: segment pairs dealt by cost (longest processing time first)
b = syn.make_config(1, scale=0.2)
T21, T12, mine, cnt = D.local_icp_distributed(b, compute=compute, gather=True)
R21, R12, Rc = compute(b)
assert np.array_equal(T21, R21) and np.array_equal(T12, R12), "gathered maps differ from the single-process result"
assert np.array_equal(cnt, Rc), "gathered score counts differ"
assert 0 < len(mine) < b.n_segments
# (2) a growth series: whole frame pairs in contiguous blocks, every rank keeps the frames its pairs share
ser = syn.make_series(1, n_frames=4, scale=0.07)
sb = ser.as_batch()
T21, T12, mine, cnt = D.local_icp_distributed(sb, compute=compute, gather=True)
R21, R12, Rc = compute(sb)
assert np.array_equal(T21, R21) and np.array_equal(T12, R12) and np.array_equal(cnt, Rc)
fps = np.unique(sb.frame_pair[mine])
assert np.array_equal(fps, np.arange(fps.min(), fps.max() + 1)), "frame pairs of a rank must be contiguous"
# (3) the packed gather on raw tensors: int32 maps + int64 counts in ONE all_gather_into_tensor
plan = D.make_shard_plan(sb.off1, sb.off2, sb.lm_off, 2, sb.frame_pair, exchange=True)
sub = sb.take(plan.seg_ids[dist.get_rank()])
l21, l12, lc = compute(sub)
g21, g12, gc = D.gather_results(plan, torch.as_tensor(l21), torch.as_tensor(l12), torch.as_tensor(lc))
assert g21.dtype == torch.int32 and np.array_equal(g21.numpy(), R21) and np.array_equal(g12.numpy(), R12) and np.array_equal(gc.numpy(), Rc)
print("rank", dist.get_rank(), "ok", len(mine))
dist.destroy_process_group()
'''


def test_world_size_2_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, LIP_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_result_files_are_what_the_consumers_read(tmp_path, syn):
    """the on-disk format of p2p_matching_seg_icp.py:210-242 round-trips through the reader
    logic of interpolation/interpolation_pcd.py:15-55 (np.loadtxt of {day}_{i}.csv, T21_{i}.csv, T12_{i}.csv)."""
    B = importlib.import_module("4d_plant_analysis_b200.batch")
    b = syn.make_config(1, scale=0.1)
    rng = np.random.default_rng(0)
    T21s = [rng.integers(0, b.off1[s + 1] - b.off1[s], b.off2[s + 1] - b.off2[s]) for s in range(b.n_segments)]
    T12s = [rng.integers(0, b.off2[s + 1] - b.off2[s], b.off1[s + 1] - b.off1[s]) for s in range(b.n_segments)]
    res = B.SegmentMatchResult(T21s, T12s, rng.random((b.n_segments, 3)), np.diff(b.off1).astype(float), b)
    B.save_segment_results(res, str(tmp_path), "03-22_AM", "03-23_AM")
    i = 0
    while os.path.exists(tmp_path / f"03-22_AM_{i}.csv"):                  # the reader's loop
        x1 = np.loadtxt(tmp_path / f"03-22_AM_{i}.csv"); x2 = np.loadtxt(tmp_path / f"03-23_AM_{i}.csv")
        t21 = np.loadtxt(tmp_path / f"T21_{i}.csv"); t12 = np.loadtxt(tmp_path / f"T12_{i}.csv")
        np.testing.assert_array_equal(x1, b.segment(i)[0]); np.testing.assert_array_equal(x2, b.segment(i)[1])
        np.testing.assert_array_equal(t21.astype(int), T21s[i]); np.testing.assert_array_equal(t12.astype(int), T12s[i])
        i += 1
    assert i == b.n_segments
    np.testing.assert_allclose(np.loadtxt(tmp_path / "score.csv"), res.scores)
    np.testing.assert_allclose(res.score, np.average(res.scores, axis=0, weights=np.diff(b.off1)))


def test_skeleton_landmarks_match_reference():
    """SURVEY.md section 8 f4 (host logic): the landmark sampling of segment_matching_ml.py:210-239 against values produced
    by those very source lines (tests/golden/make_golden_next.py)."""
    from conftest import load_golden
    S = importlib.import_module("4d_plant_analysis_b200.segment_matching_utils")
    nx = load_golden("next")
    for ci in nx["lm_cases"]:
        vox, method = nx[f"lm_{ci}_par"]
        lm1, lm2 = S.skeleton_landmarks(nx[f"lm_{ci}_skel1"], nx[f"lm_{ci}_skel2"], vox, "end" if method == 0 else "uniform")
        np.testing.assert_array_equal(lm1, nx[f"lm_{ci}_lm1"])
        np.testing.assert_array_equal(lm2, nx[f"lm_{ci}_lm2"])


def test_install_patches_every_reference_module(monkeypatch):
    """install() replaces the functions of the path (and of the rows either side of it) in the reference's modules by name,
    wherever the reference resolves them as module globals.  Stub modules stand in for the reference."""
    import types
    L = importlib.import_module("4d_plant_analysis_b200")
    names = {
        "registration.registration_icp": ["get_best_fit_transform", "get_NN_in_target", "my_icp"],
        "p2p_matching_in_organ.p2p_matching_seg_icp": ["get_triangle_angles", "get_p2p_mapping_from_segments_using_icp", "my_icp",
                                                       "evaluate_p2p_matching", "double_direction_refinement"],
        "p2p_matching_in_organ.p2p_matching_evaluation": ["evaluate_p2p_matching"],
        "p2p_matching_in_organ.p2p_matching_fm": ["double_direction_refinement"],
        "interpolation.interpolation_pcd": ["interpolation_segment_level"],
    }
    for pkg in ("registration", "p2p_matching_in_organ", "interpolation"):
        monkeypatch.setitem(sys.modules, pkg, types.ModuleType(pkg))
    mods = {}
    for name, fns in names.items():
        m = types.ModuleType(name)
        for f in fns:
            setattr(m, f, lambda *a, **k: "reference")
        monkeypatch.setitem(sys.modules, name, m)
        mods[name] = m
    done = L.install()
    for name, fns in names.items():
        for f in fns:
            fn = getattr(mods[name], f)
            assert fn.__module__.startswith("4d_plant_analysis_b200"), (name, f)
            assert (name, f) in done


_REAL_REF = r'''
import importlib, inspect, os, sys
root = os.environ["LIP_ROOT"]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests", "golden"))
import reference_loader
ref = reference_loader.load()                      # the UNMODIFIED reference modules, imported under the stub recipe
names = {ref.RI: ("get_best_fit_transform", "get_NN_in_target", "my_icp"),
         ref.SI: ("get_triangle_angles", "get_p2p_mapping_from_segments_using_icp", "my_icp", "double_direction_refinement",
                  "evaluate_p2p_matching"),
         ref.FM: ("double_direction_refinement",), ref.EV: ("evaluate_p2p_matching",)}
before = {(m.__name__, n): (getattr(m, n), inspect.signature(getattr(m, n))) for m, ns in names.items() for n in ns}
L = importlib.import_module("4d_plant_analysis_b200")
done = L.install()
for (mod, n), (fn0, sig0) in before.items():
    fn1 = getattr(sys.modules[mod], n)
    assert fn1 is not fn0, f"{mod}.{n} was not replaced"
    assert fn1.__module__.startswith("4d_plant_analysis_b200"), (mod, n, fn1.__module__)
    assert str(inspect.signature(fn1)) == str(sig0), f"{mod}.{n}: {inspect.signature(fn1)} != {sig0}"
    assert (mod, n) in done
print("real reference patched:", len(before), "bindings")
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree is only present in the authoring container")
def test_install_patches_the_real_reference_modules(tmp_path):
    """install() against the REAL reference modules (not stubs): every hot function, and the names the
    driver module imported from them, is re-bound to this package with a signature string-identical to the reference's.  Runs in
    a subprocess (the stub recipe edits sys.modules / sys.path)."""
    script = tmp_path / "real_ref.py"
    script.write_text(_REAL_REF)
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=dict(os.environ, LIP_ROOT=ROOT), timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "real reference patched: 10 bindings" in r.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the base
    contract's keys, `impl`, a cpu_baseline that describes the run and an e2e that repeats the value"""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "local_icp point correspondences/sec" and line["unit"] == "correspondences/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
