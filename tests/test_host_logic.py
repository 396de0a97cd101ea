"""Host-side scheduling rules that need no GPU: when the RGB and the depth modality call are worth overlapping."""
import pytest


def test_modality_overlap_rule_on_a_b200_sized_grid():
    from cosnet_b200.coattention import modality_overlap_pays
    pays = lambda n, h, w, passes=2: modality_overlap_pays(n, h, w, passes, clusters=74)
    # items per call = passes * n * ceil(h w / 256); overlap iff 2 * ceil(items / 74) > ceil(2 items / 74)
    assert pays(1, 60, 60)             # 30 items: 2 waves -> 1
    assert not pays(2, 60, 60)         # 60: 2 -> 2
    assert not pays(4, 60, 60)         # 120: 4 -> 4
    assert pays(8, 60, 60)             # 240: 8 -> 7   (cfg 5 forward)
    assert not pays(32, 60, 60)        # 960: 26 -> 26 (headline batch)
    assert pays(2, 61, 107)            # 104: 4 -> 3   (cfg 3 on 8 GPUs)
    assert not pays(4, 61, 107)        # 208: 6 -> 6
    assert pays(16, 61, 107)           # 832: 24 -> 23
    assert pays(5, 61, 81, 1)          # test.py: 1 query x 5 references, 100 frame-A items: 4 -> 3
    assert not pays(40, 61, 81, 1)     # cfg 4 batch: 800 items: 22 -> 22
    # a grid that is already a whole number of waves never pays
    for n in range(1, 20):
        items = 2 * n * 15
        assert pays(n, 60, 60) == (2 * -(-items // 74) > -(-2 * items // 74))


def test_run_modalities_without_overlap_is_a_plain_sequence():
    from cosnet_b200.coattention import run_modalities
    order = []
    r, d = run_modalities(lambda: order.append("rgb") or "R", lambda: order.append("depth") or "D", (), overlap=False)
    assert (r, d) == ("R", "D") and order == ["rgb", "depth"]
