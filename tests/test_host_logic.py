"""CPU tests of the host-side mirror of the reference interface (no GPU needed)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden


def cfg_dict(**over):
    cfg = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    cfg.update(over)
    return cfg


def test_config_loader_derived_names_match_reference_fixture():
    from geobo_b200 import config_loader as cl
    from oracle import numpy_oracle as o
    cfg = cfg_dict()
    cl.load_settings(cfg, make_outpath=False)
    ref = o.make_config(cfg)
    for k in ("xLcube", "yLcube", "zmin", "c_MILLIGALS_UNITS", "xvoxsize", "yvoxsize", "zvoxsize", "Nsensor"):
        assert getattr(cl, k) == ref[k], k
    assert np.array_equal(cl.magneticField, ref.magneticField)
    assert cl.xNcube == 8 and cl.kernelfunc == cfg["kernelfunc"]
    # star-import surface
    ns = {}
    exec("from geobo_b200.config_loader import *", ns)
    assert ns["xvoxsize"] == ref.xvoxsize and "gp_coeff" in ns


def test_config_loader_reads_yaml_from_argv(tmp_path):
    import yaml
    cfg = cfg_dict(outpath=str(tmp_path / "out") + os.sep, inpath=str(tmp_path) + os.sep, FNAME_drilldata="d.csv",
                   FNAME_gravsurvey="g.tif", FNAME_magsurvey="m.tif")
    y = tmp_path / "settings.yaml"
    y.write_text(yaml.safe_dump(cfg))
    code = ("import sys; sys.argv=['main.py', %r]; sys.path.insert(0, %r);"
            "from geobo_b200.config_loader import *; import os;"
            "print(xvoxsize, Nsensor, os.path.isdir(outpath), fname_gravsurvey)") % (str(y), ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.split()
    assert float(out[0]) == 3050 / 8 and int(out[1]) == 48 and out[2] == "True" and out[3].endswith("g.tif")


def test_unconfigured_inversion_raises_clear_error():
    code = ("import sys; sys.argv=['x']; sys.path.insert(0, %r); import os; os.chdir('/tmp');"
            "from geobo_b200 import inversion\n"
            "try:\n inversion.Inversion()\nexcept RuntimeError as e:\n print('OK', e)") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert out.startswith("OK") and "settings not loaded" in out


def test_dedup_lengthscales_quirk_q1():
    from geobo_b200.kernels import dedup_lengthscales
    a = np.array([244.0, 244.0, 244.0])
    dedup_lengthscales(a)
    assert np.allclose(a, [244.0, 248.88, 244.0])       # second test rewrites element 1, element 2 untouched
    b = np.array([500, 500, 500])                        # int ndarray truncates
    dedup_lengthscales(b)
    assert b.tolist() == [500, 510, 500]
    c = np.array([1.0, 2.0, 2.0])
    dedup_lengthscales(c)
    assert np.allclose(c, [1.0, 2.0, 2.02])
    d = np.array([1.0, 2.0, 3.0])
    dedup_lengthscales(d)
    assert d.tolist() == [1.0, 2.0, 3.0]


def test_geometry_matches_reference_fixture():
    from geobo_b200 import config_loader as cl, inversion
    s = load_golden("sens_8x6x5.npz")
    cl.load_settings(json.loads(str(s["cfg"])), make_outpath=False)
    inv = inversion.Inversion()
    vp = inv.create_cubegeometry()
    assert np.array_equal(vp, s["voxelpos"]) and np.array_equal(inv.Edges, s["Edges"])
    assert inv.xxx.shape == (6, 8, 5)
    assert np.allclose(inv.gp_length, 2 * 3050 / 8)       # Q6: xvoxsize for all three


def test_downsample_survey_matches_the_reference_driver():
    """tests/golden/survey_zoom.npz: the raw survey images of the reference's example 2 (61 x 39 GeoTIFFs) and what its own
    read_surveydata handed to Inversion.cubing (the `grav`, `mag`, `sensor_locations` of example2.npz, captured from the
    unmodified driver)."""
    from geobo_b200 import config_loader as cl, utils
    g = load_golden("survey_zoom.npz")
    cl.load_settings(json.loads(str(g["cfg"])), make_outpath=False)
    grav, mag, loc = utils.downsample_survey(g["grav_img"], g["mag_img"])
    assert np.array_equal(grav, g["grav"]) and np.array_equal(mag, g["mag"]) and np.array_equal(loc, g["sensor_locations"])


def test_shard_columns_partition():
    from geobo_b200.dist import shard_columns
    for n, world in [(6400, 1), (6400, 2), (32768, 8), (442368, 8), (1000, 3), (131072, 4)]:
        edges = [shard_columns(n, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a0 % 128 == 0 and a1 > a0
    with pytest.raises(ValueError):
        shard_columns(200, 4, 3)


def test_shard_bounds_balance_the_cumulative_weight():
    """Weighted contiguous shards (culled projection work per voxel row): aligned, exhaustive, every rank non-empty, and the edge
    ranks -- whose rows see fewer rows inside the covariance's reach -- get more columns."""
    from geobo_b200.dist import shard_bounds, shard_columns
    assert shard_bounds(6400, 3) == [shard_columns(6400, 3, r) for r in range(3)]
    yN, XZ, reach = 96, 96 * 48, 23
    iy = np.arange(yN)
    nrows = np.minimum(yN - 1, iy + reach) - np.maximum(0, iy - reach) + 1
    w = 0.75 * nrows / nrows.max() + 0.25
    col_w = np.repeat(w, XZ)
    for world in (2, 3, 4, 8):
        b = shard_bounds(yN * XZ, world, weights=w)
        assert b[0][0] == 0 and b[-1][1] == yN * XZ and all(a[1] == c[0] for a, c in zip(b, b[1:]))
        assert all(c0 % 128 == 0 and c1 > c0 for c0, c1 in b)
        loads = np.array([col_w[c0:c1].sum() for c0, c1 in b])
        assert loads.max() / loads.min() < 1.02
        if world >= 4:
            assert b[0][1] - b[0][0] > b[1][1] - b[1][0]
    assert shard_bounds(512, 4, weights=[100, 1, 1, 1]) == [(0, 128), (128, 256), (256, 384), (384, 512)]   # never an empty rank
    with pytest.raises(ValueError):
        shard_bounds(256, 4, weights=[1, 1])


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from geobo_b200 import dist
    dist.init_from_env(ctx=None)
    n = 1000
    c0, c1 = dist.shard_columns(n, world, rank)
    full_truth = np.arange(3 * n, dtype=float).reshape(3, n)
    got = dist.allgather_columns(full_truth[:, c0:c1], n)
    # unequal (work-balanced) shards: the gather pads to the largest one
    wb = dist.shard_bounds(n, world, weights=[1.0, 1.0, 1.0, 3.0, 3.0, 3.0, 3.0, 3.0])
    assert wb[0][1] - wb[0][0] != wb[1][1] - wb[1][0]
    got_w = dist.allgather_columns(full_truth[:, wb[rank][0]:wb[rank][1]], n, bounds=wb)
    assert np.array_equal(got_w, full_truth)
    mx = dist.max_over_ranks(10.0 + rank)
    uid = dist.broadcast_bytes(b"id-from-rank-0" if rank == 0 else None)
    dist.barrier()
    # the same gather / max through torch.distributed's gloo backend (test infrastructure only: the package itself has no torch)
    import torch
    import torch.distributed as td
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    per = dist.shard_columns(n, world, 0)[1]
    buf = np.zeros((3, per))
    buf[:, :c1 - c0] = full_truth[:, c0:c1]
    outs = [torch.zeros((3, per), dtype=torch.float64) for _ in range(world)]
    td.all_gather(outs, torch.from_numpy(buf))
    gl = np.empty((3, n))
    for r in range(world):
        a0, a1 = dist.shard_columns(n, world, r)
        gl[:, a0:a1] = outs[r].numpy()[:, :a1 - a0]
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    td.destroy_process_group()
    dist.shutdown()
    q.put((rank, bool(np.array_equal(got, full_truth)) and bool(np.array_equal(got, gl)) and uid == b"id-from-rank-0", mx, float(t[0])))


def test_two_rank_gloo_gather_and_max():
    """World-size-2 run of the host-side shard / gather / timing logic of the multi-GPU path: the package's own TCP control
    plane (geobo_b200/dist.py; no torch in the product) next to a gloo process group doing the same gather and max."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True, 11.0, 11.0), (1, True, 11.0, 11.0)]


def test_package_has_no_torch_import():
    import re
    pkg = os.path.join(ROOT, "geobo_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(import|from)\s+torch", src, re.M), name
