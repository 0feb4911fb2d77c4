"""Size-independent properties of the integer / index kernels at BASELINE.json's full sizes (1024 x 1024 masks,
hundreds to thousands of prompts), where a CPU oracle run would take minutes: consistency between the two K-POST
passes, NMS idempotence, RLE encode -> decode round trips, connected-component cleanup idempotence / monotonicity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fixtures  # noqa: E402

DEV = "cuda"


def ops():
    from crowdsam_b200 import ops as o

    return o


def test_mask_post_stats_and_write_agree_full_size():
    """counts[:,2] = number of mask pixels, boxes = inclusive extent of the written mask, and the three counts are
    ordered, for 256 prompts at 1024 x 1024 (blob logits + noise)."""
    from crowdsam_b200 import amg

    o = ops()
    P = 256
    low, _, _ = fixtures.blob_logits(P, seed=11)
    low = low.to(DEV)
    sel = torch.randint(0, 4, (P,), generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    counts, boxes = o.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
    masks, _ = o.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
    assert masks.shape == (P, 1024, 1024) and masks.dtype == torch.bool
    assert torch.equal(counts[:, 2].long(), masks.flatten(1).sum(1))
    assert bool((counts[:, 0] <= counts[:, 2]).all()) and bool((counts[:, 2] <= counts[:, 1]).all())
    assert torch.equal(boxes.long(), amg.batched_mask_to_box(masks).long())
    # keep-list variant writes exactly the selected rows
    keep = torch.arange(0, P, 3, dtype=torch.int32, device=DEV)
    sub, _ = o.mask_post_write(low, sel, keep, (1024, 1024), (1024, 1024), 0.0)
    assert torch.equal(sub, masks[keep.long()])


@pytest.mark.parametrize("n", [1024, 4096])
def test_nms_idempotent_and_sorted(n):
    """NMS of the kept set keeps everything; kept indices come in non-increasing score order; no kept pair overlaps
    above the threshold."""
    o = ops()
    b, s = fixtures.random_boxes(n, seed=n, binary_scores=False)
    boxes, scores = torch.as_tensor(b).to(DEV), torch.as_tensor(s).to(DEV)
    keep = o.box_nms(boxes, scores, 0.65)
    ks = scores[keep]
    assert bool((ks[:-1] >= ks[1:]).all())
    again = o.box_nms(boxes[keep], ks, 0.65)
    assert torch.equal(again, torch.arange(len(keep), device=DEV))
    kb = boxes[keep][:512]
    x1 = torch.maximum(kb[:, None, 0], kb[None, :, 0]); y1 = torch.maximum(kb[:, None, 1], kb[None, :, 1])
    x2 = torch.minimum(kb[:, None, 2], kb[None, :, 2]); y2 = torch.minimum(kb[:, None, 3], kb[None, :, 3])
    inter = (x2 - x1).clamp(min=0) * (y2 - y1).clamp(min=0)
    area = (kb[:, 2] - kb[:, 0]) * (kb[:, 3] - kb[:, 1])
    iou = inter / (area[:, None] + area[None, :] - inter)
    iou.fill_diagonal_(0)
    assert float(iou.nan_to_num(0).max()) <= 0.65


def test_rle_round_trip_full_size():
    from crowdsam_b200 import amg

    o = ops()
    g = torch.Generator().manual_seed(5)
    m = torch.nn.functional.interpolate(torch.randn(24, 1, 32, 32, generator=g), (1024, 1024), mode="bilinear")[:, 0] > 0.2
    m ^= torch.rand(24, 1024, 1024, generator=g) < 0.002
    m[0] = False
    m[1] = True
    runs = o.rle_encode(m.to(DEV))
    for i in range(m.shape[0]):
        assert int(np.sum(runs[i])) == 1024 * 1024
        back = amg.rle_to_mask({"size": [1024, 1024], "counts": runs[i]})
        assert np.array_equal(back, m[i].numpy())
        s = amg._coco_string(runs[i])
        assert np.array_equal(np.asarray(_coco_decode(s)), np.asarray(runs[i], dtype=np.int64))


def _coco_decode(s):
    """Inverse of the COCO run-length string (maskApi.c rleFrString)."""
    counts, p, b = [], 0, s.encode("ascii")
    while p < len(b):
        x, k, more = 0, 0, True
        while more:
            c = b[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def test_small_region_cleanup_idempotent_monotone_full_size():
    o = ops()
    g = torch.Generator().manual_seed(9)
    m = torch.nn.functional.interpolate(torch.randn(48, 1, 48, 48, generator=g), (1024, 1024), mode="bilinear")[:, 0] > 0.3
    m ^= torch.rand(48, 1024, 1024, generator=g) < 0.004
    m0 = m.to(torch.uint8).to(DEV).contiguous()
    a = m0.clone()
    c = o.remove_small_regions(a, 100, "holes")
    assert bool((a >= m0).all()) and bool(c.bool().any())            # holes are only ever filled
    a2 = a.clone()
    c2 = o.remove_small_regions(a2, 100, "holes")
    assert torch.equal(a2, a) and not bool(c2.bool().any())            # nothing small is left
    b = a.clone()
    o.remove_small_regions(b, 100, "islands")
    assert bool((b <= a).all())                                        # islands are only ever removed
    b2 = b.clone()
    c4 = o.remove_small_regions(b2, 100, "islands")
    assert torch.equal(b2, b) and not bool(c4.bool().any())
