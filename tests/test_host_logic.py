"""CPU checks of host-side logic that the GPU tests rely on but that needs no device."""
import numpy as np
import torch

from crowdsam_b200 import amg


def _coco_string_scalar(counts):
    """Straight restatement of maskApi.c rleToString (published COCO API algorithm), one value at a time."""
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def test_coco_string_vectorised_matches_scalar():
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 3, 4, 7, 1000, 5000):
        for hi in (15, 70000, 1 << 20):
            counts = rng.integers(0, hi, size=n)
            assert amg._coco_string(counts) == _coco_string_scalar(counts), (n, hi)


def test_rle_round_trip():
    rng = np.random.default_rng(3)
    m = rng.random((37, 53)) < 0.4
    # column-major runs, first run counts zeros (amg.py:107-135)
    flat = m.T.reshape(-1)
    change = np.flatnonzero(np.diff(flat.astype(np.int8))) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = list(np.diff(bounds))
    if flat[0]:
        counts = [0] + counts
    rle = {"size": [37, 53], "counts": counts}
    assert np.array_equal(amg.rle_to_mask(rle), m)
    assert amg.area_from_rle(rle) == int(m.sum())


def test_coco_string_batched_library_call_matches_scalar():
    """csam_coco_rle_strings (host C, one call for all masks of an image) against the scalar restatement of
    maskApi.c rleToString, incl. empty run lists, single runs and values needing up to 5 characters."""
    rng = np.random.default_rng(11)
    rles = []
    for n, hi in ((0, 10), (1, 10), (2, 1 << 20), (3, 70000), (4, 15), (500, 3000), (5000, 1 << 20), (7, 1 << 20)):
        rles.append({"size": [1024, 1024], "counts": rng.integers(0, hi, size=n)})
    rles.append({"size": [4, 4], "counts": np.array([16])})
    rles.append({"size": [4, 4], "counts": [0, 16]})
    got = amg.coco_encode_rles(rles)
    assert [g["counts"] for g in got] == [_coco_string_scalar(list(r["counts"])) for r in rles]
    assert [g["size"] for g in got] == [list(r["size"]) for r in rles]
    assert amg.coco_encode_rles([]) == []


def test_tensors_to_numpy_packed_equals_per_tensor_copies():
    """MaskData.to_numpy reads all device columns back in one transfer: same dtypes, shapes, values (any device)."""
    from crowdsam_b200.amg import tensors_to_numpy_packed

    ts = [torch.randn(5, 4), torch.arange(7), torch.zeros(0, 4), torch.tensor([True, False, True]),
          torch.randn(3).double(), torch.arange(3, dtype=torch.int32), torch.randn(0), torch.randn(2, 3, 5)[:, :, 1]]
    for t, a in zip(ts, tensors_to_numpy_packed(ts)):
        r = t.numpy()
        assert a.dtype == r.dtype and a.shape == r.shape and np.array_equal(a, r)
    assert tensors_to_numpy_packed([]) == []
    assert tensors_to_numpy_packed([torch.zeros(0)])[0].shape == (0,)


def test_interleave_two_streams_protocol(monkeypatch):
    """engine.interleave_two_streams (host logic of the two-stream encoders) with fake streams: the side generator is only
    ever resumed inside the side-stream context, the main one outside, block by block in alternation; the side stream first
    waits for the main stream and the main stream waits for the side stream at the end; both results come back."""
    import contextlib

    from crowdsam_b200 import engine

    log = []

    class FakeStream:
        def __init__(self, name="side", **kw):
            self.name = name

        def wait_stream(self, other):
            log.append(("wait", self.name, other.name))

    main = FakeStream("main")
    current = [main]

    @contextlib.contextmanager
    def fake_ctx(s):
        current.append(s)
        try:
            yield
        finally:
            current.pop()

    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: main)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", fake_ctx)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(engine, "_side_streams", {})

    def gen(tag, blocks):
        log.append((tag, "prologue", current[-1].name))
        for b in range(blocks):
            yield
            log.append((tag, b, current[-1].name))
        return tag + "-result"

    res = engine.interleave_two_streams(gen("sam", 3), gen("dino", 5), "cuda:0")
    assert res == ("sam-result", "dino-result")
    assert log[0] == ("wait", "side", "main") and log[-1] == ("wait", "main", "side")
    steps = [e for e in log if e[0] in ("sam", "dino")]
    assert all(e[2] == ("main" if e[0] == "sam" else "side") for e in steps)
    assert [e[1] for e in steps if e[0] == "sam"] == ["prologue", 0, 1, 2]
    assert [e[1] for e in steps if e[0] == "dino"] == ["prologue", 0, 1, 2, 3, 4]
    # alternation while both are live: dino k, sam k, dino k+1, ...
    order = [e[0] for e in steps[:8]]
    assert order == ["dino", "sam"] * 4
    # run_steps: the single-stream driver of the same generators
    assert engine.run_steps(gen("sam", 2)) == "sam-result"
