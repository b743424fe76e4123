"""Host-side logic that needs no GPU: parameter planner of the C++ graph builder vs the oracle's layer list,
checkpoint byte layout, batch sharding."""
import ctypes
import os

import numpy as np
import pytest


@pytest.mark.parametrize("kw,okw", [
    (dict(), dict()),
    (dict(H=128, W=128, channel_mult=(1, 1, 2, 3, 4), att_start_level=3),
     dict(H=128, W=128, channel_mult=(1, 1, 2, 3, 4), attn_start_level=3)),
    (dict(H=32, W=32, channel_mult=(1, 2), att_start_level=1), dict(H=32, W=32, channel_mult=(1, 2), attn_start_level=1)),
    (dict(C_model=128, channel_mult=(1, 2, 2), att_start_level=2, n_res_blocks=1),
     dict(model_channels=128, channel_mult=(1, 2, 2), attn_start_level=2, num_res_blocks=1)),
])
def test_builder_param_count_matches_oracle(ub, oracle, kw, okw):
    """ub_num_params runs the C++ graph builder in counting mode (no CUDA call)."""
    cfg = ub.default_config(**kw)
    assert ub.num_params(cfg) == oracle.num_params(oracle.UNetConfig(**okw))


def test_default_config_is_the_reference_model(ub):
    cfg = ub.default_config()
    assert (cfg.B, cfg.C_in, cfg.C_model, cfg.C_out, cfg.H, cfg.W, cfg.max_period) == (32, 3, 64, 3, 64, 64, 1000)
    assert list(cfg.channel_mult)[:4] == [1, 2, 3, 4] and cfg.n_levels == 4 and cfg.att_start_level == 2
    assert ub.num_params(cfg) == 20494211


def test_bad_configs_are_rejected_without_gpu(ub):
    assert ub.num_params(ub.default_config(head_size=16)) == 0
    assert ub.num_params(ub.default_config(C_model=48)) == 0
    assert ub.num_params(ub.default_config(H=60)) == 0


def test_checkpoint_header_reader(ub, oracle, tmp_path):
    O = oracle
    cfg = O.UNetConfig(channel_mult=(1, 2), attn_start_level=1, H=32, W=32)
    flat = np.arange(O.num_params(cfg), dtype=np.float32)
    path = str(tmp_path / "m.bin")
    O.write_model_bin(path, cfg, flat, B=8)
    c = ub.default_config()
    rc = ub.lib().ub_read_checkpoint_header(path.encode(), ctypes.byref(c))
    assert rc == 0 and (c.B, c.C_in, c.C_model, c.C_out, c.H, c.W, c.max_period) == (8, 3, 64, 3, 32, 32, 1000)
    with open(path, "r+b") as f:
        f.write(b"\0\0\0\0")
    assert ub.lib().ub_read_checkpoint_header(path.encode(), ctypes.byref(c)) != 0
    assert b"magic" in ub.lib().ub_last_error()


def test_data_bin_format(oracle, tmp_path):
    """prepare_data.py:20-38 layout."""
    imgs = np.random.default_rng(0).uniform(-1, 1, (5, 3, 8, 8)).astype(np.float32)
    p = str(tmp_path / "d.bin")
    oracle.write_data_bin(p, imgs)
    hdr = np.fromfile(p, dtype=np.int32, count=256)
    assert list(hdr[:5]) == [20240620, 5, 3, 8, 8]
    np.testing.assert_array_equal(np.fromfile(p, dtype=np.float32, offset=1024).reshape(imgs.shape), imgs)


def test_shard_batch(ub):
    assert [ub.shard_batch(256, r, 8) for r in (0, 3, 7)] == [(0, 32), (96, 128), (224, 256)]
    with pytest.raises(ValueError):
        ub.shard_batch(10, 0, 4)


def test_dataloader_batches_wrap_and_rank_striding(ub, oracle, tmp_path):
    """ub_dataloader_* vs the semantics of DataLoader / dataloader_next_batch (train_unet.cu:3035-3099): sequential
    batches of B, wrap to the start when fewer than B images remain; rank r of R reads global batches r, r+R, ..."""
    rng = np.random.default_rng(3)
    imgs = rng.uniform(-1, 1, (11, 3, 8, 8)).astype(np.float32)       # 11 images, B = 4 -> 2 full batches per epoch
    path = str(tmp_path / "data.bin")
    oracle.write_data_bin(path, imgs)
    dl = ub.DataLoader(path, B=4)
    assert (dl.n_imgs, dl.shape, dl.batches_per_epoch) == (11, (3, 8, 8), 2)
    for k in range(5):                                                 # 0, 1, wrap -> 0, 1, 0
        np.testing.assert_array_equal(dl.next(), imgs[(k % 2) * 4:(k % 2) * 4 + 4])
    dl.reset()
    np.testing.assert_array_equal(dl.next(), imgs[0:4])
    # three buffers: the batch of one call stays intact through the next call (current + announced next batch, the
    # pattern ub_trainer_set_next_batch needs), and is recycled by the call after that
    import ctypes as C
    n = 4 * 3 * 8 * 8
    view = lambda p: np.frombuffer((C.c_float * n).from_address(p), dtype=np.float32).reshape(4, 3, 8, 8)
    p1 = dl.next_ptr()                                                 # batch 1
    keep1 = view(p1).copy()
    p2 = dl.next_ptr()                                                 # batch 0 (wrapped), another buffer
    assert p2 != p1
    np.testing.assert_array_equal(view(p1), keep1)
    np.testing.assert_array_equal(view(p2), imgs[0:4])
    p3 = dl.next_ptr()
    assert p3 not in (p1, p2)
    np.testing.assert_array_equal(view(p2), imgs[0:4])
    dl.close()
    imgs = rng.uniform(-1, 1, (24, 1, 4, 4)).astype(np.float32)        # 6 batches of 4, two ranks
    oracle.write_data_bin(path, imgs)
    r0, r1 = ub.DataLoader(path, B=4, rank=0, world=2), ub.DataLoader(path, B=4, rank=1, world=2)
    for k in range(4):                                                 # rank 0: 0,2,4,0   rank 1: 1,3,5,1
        np.testing.assert_array_equal(r0.next(), imgs[((2 * k) % 6) * 4:((2 * k) % 6) * 4 + 4])
        np.testing.assert_array_equal(r1.next(), imgs[((2 * k + 1) % 6) * 4:((2 * k + 1) % 6) * 4 + 4])
    r0.close(), r1.close()
    with pytest.raises(ub.UbError):
        ub.DataLoader(str(tmp_path / "missing.bin"), B=4)
    (tmp_path / "bad.bin").write_bytes(b"\x00" * 2048)
    with pytest.raises(ub.UbError):
        ub.DataLoader(str(tmp_path / "bad.bin"), B=4)
