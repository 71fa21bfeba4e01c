"""GPU (-m gpu), needs >= 2 GPUs (skipped otherwise; `gpurun --gpus 2`): the multi-process path -- one rank per GPU under
torch.distributed.run -- through the library's own NCCL calls: vp_nccl_init, the sharded opacity build (each rank
sweeps 1/G of the table, ncclAllGather) and vp_reduce_nccl of the sample-sharded sums, against the one-GPU result."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["VP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["VP_ROOT"], "tests"))
import cuda_volpath_b200 as vp
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r = vp.Renderer(local)
env, sd, sp = vp.default_sunsky()
r.generate_cloud(120, 80, 144, seed=2, bounds=vp.BOUNDS_CELL)
r.set_texture_filter_mode(True); r.init_envmap(env); r.set_sun(sd, sp); r.copy_inv_view_matrix(vp.inv_view_matrix())
vp.init_nccl_from_torch(r)
r.precompute_opacity(sd, sharded=True)
tab = r.opacity_fast()
P = vp.default_param(128, 96); P.density = 1500.0
acc = torch.zeros(96, 128, 4, device="cuda")
first, count, stride = vp.frames_for_rank(11, 37, rank, world)
s = torch.cuda.current_stream().cuda_stream
r.render_kernel(acc.data_ptr(), first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=s)
# the same sums once more through peer memory (CUDA IPC): plain cudaMalloc blocks, handles through the TCP store
blk = r.dev_alloc(128 * 96 * 16)
r.render_kernel(blk, first, P, mode=vp.MODE_FAST, n_frames=count, frame_stride=stride, stream=s)
torch.cuda.synchronize()
store = dist.distributed_c10d._get_default_store()
store.set("h%d" % rank, r.ipc_export(blk))
r.reduce_nccl(acc.data_ptr(), acc.data_ptr() if rank == 0 else None, 128 * 96, root=0, stream=s)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    r.reduce_ipc(blk, [bytes(store.get("h%d" % q)) for q in range(1, world)], 128 * 96, stream=s)
    torch.cuda.synchronize()
    ipc = np.empty((96, 128, 4), np.float32)
    assert r.L.vp_dev_to_host(ipc.ctypes.data, blk, ipc.nbytes) == 0
    np.save(os.environ["VP_OUT"] + "_ipc.npy", ipc)
    np.save(os.environ["VP_OUT"] + "_img.npy", acc.cpu().numpy()); np.save(os.environ["VP_OUT"] + "_tab.npy", tab)
dist.barrier()
dist.barrier(); r.close(); dist.destroy_process_group()
'''


def test_two_ranks_sharded_opacity_and_nccl_reduce_equal_one_gpu(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import cuda_volpath_b200 as vp

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "two")
    env = dict(os.environ, VP_ROOT=ROOT, VP_OUT=out)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    img2, tab2 = np.load(out + "_img.npy"), np.load(out + "_tab.npy")
    r = vp.Renderer(0)
    env_map, sd, sp = vp.default_sunsky()
    r.generate_cloud(120, 80, 144, seed=2, bounds=vp.BOUNDS_CELL)
    r.set_texture_filter_mode(True)
    r.init_envmap(env_map)
    r.set_sun(sd, sp)
    r.copy_inv_view_matrix(vp.inv_view_matrix())
    r.precompute_opacity(sd)
    assert np.array_equal(tab2, r.opacity_fast())           # the gathered table is the table
    P = vp.default_param(128, 96)
    P.density = 1500.0
    img1 = r.render(P, 11, 37, mode=vp.MODE_FAST)
    r.close()
    assert img1[..., 3].max() / 37 > 20                     # deep paths: the opacity-table branch is exercised
    assert np.array_equal(img1[..., 3], img2[..., 3])       # every (pixel, frame) exactly once, same paths
    assert np.allclose(img1[..., :3], img2[..., :3], rtol=1e-5, atol=1e-6)
    ipc = np.load(out + "_ipc.npy")                         # vp_reduce_ipc: the same sums over peer memory
    assert np.array_equal(ipc[..., 3], img2[..., 3]) and np.allclose(ipc[..., :3], img2[..., :3], rtol=1e-5, atol=1e-6)
