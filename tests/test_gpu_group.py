"""pm_group_*: the multi-GPU entry points of the C ABI (one host thread, N GPUs of one box, NCCL inside the library).

Replaces -[PietRenderer initScene] + drawInMTKView: (TestApp/PietRenderer.m:203-205, :59-103) for N devices: the
scene is uploaded once and broadcast with ncclBroadcast, every device renders one contiguous strip of tile rows, and
the frame the group returns must equal the single-GPU frame byte for byte.  A group of ONE device runs everywhere
(the broadcast degenerates); the 2..8-device cases need that many GPUs and are skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def reference_frame(pm, scene, w, h):
    r = pm.PietRenderer(device=0)
    try:
        r.drawable_size_will_change(w, h)
        r.init_scene(scene)
        r.draw()
        return r.read_rgba8()
    finally:
        r.close()


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_group_frame_equals_single_gpu_frame(pm, n):
    if n > n_gpus():
        pytest.skip("needs %d GPUs" % n)
    assert pm.nccl_version() >= 21000
    w = h = 2048
    scene = pm.build_scene(pm.SCENE_TIGER, w, h)
    want = reference_frame(pm, scene, w, h)
    g = pm.PietRendererGroup(range(n))
    try:
        g.resize(w, h)
        g.set_scene(scene)
        b = g.strip_bounds()
        assert b[0] == 0 and b[-1] == (h + 15) // 16 and all(b[i] < b[i + 1] for i in range(n))
        g.render()
        stats, worst = g.sync()
        assert len(stats) == n and worst > 0 and sum(s.n_tiles for s in stats) == ((w + 15) // 16) * ((h + 15) // 16)
        assert np.array_equal(g.read_rgba8(), want)
        # the strips gathered on the last member's device over NVLink (ncclSend / ncclRecv), read back from there
        import torch
        ptr, pitch = g.gather_device(root=n - 1)
        assert pitch == 64 * ((w + 15) // 16)
        rows = ((h + 15) // 16) * 16

        class Raw:  # (a zero-copy view of the gathered device buffer)
            __cuda_array_interface__ = {"shape": (rows, pitch), "typestr": "|u1", "data": (ptr, False), "version": 2}
        with torch.cuda.device(n - 1):
            host = torch.as_tensor(Raw(), device="cuda:%d" % (n - 1)).cpu().numpy()
        assert np.array_equal(host[:h, :4 * w].reshape(h, w, 4), want)
        # a second scene through the same group (buffers and communicators are reused), frames without events
        scene2 = pm.build_scene(pm.SCENE_CARDIOID, w, h)
        g.set_scene(scene2)
        g.set_frame_events(0)
        for _ in range(3):
            g.render()
        assert np.array_equal(g.read_rgba8(), reference_frame(pm, scene2, w, h))
    finally:
        g.close()


def test_group_rejects_bad_arguments(pm):
    with pytest.raises(pm.PietMetalError):
        pm.PietRendererGroup([0, 0])          # the same device twice
    with pytest.raises(pm.PietMetalError):
        pm.PietRendererGroup([n_gpus() + 3])  # no such device
    g = pm.PietRendererGroup([0])
    try:
        with pytest.raises(pm.PietMetalError) as e:
            g.render()                         # no surface, no scene
        assert e.value.status == pm.PM_ERR_STATE
        g.resize(64, 64)
        bad = pm.build_scene(pm.SCENE_PATH_TEST, 64, 64).copy()
        bad[32:36].view(np.uint32)[0] = 1 << 30
        with pytest.raises(pm.PietMetalError) as e:
            g.set_scene(bad)
        assert e.value.status == pm.PM_ERR_SCENE_MALFORMED
    finally:
        g.close()
