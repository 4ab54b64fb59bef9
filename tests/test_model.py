"""Model compiler: GraphDef reader, layer program, arena planner, blob round trip, anchors."""
import os

import numpy as np
import pytest

from tests.conftest import REF_PB
from watsor_b200.model import (OP_ADD, OP_HEAD, OP_STEM, Model, compile_frozen_graph, ssd_anchors,
                               synthetic_ssd_mobilenet_v1)

needs_ref = pytest.mark.skipif(not os.path.isfile(REF_PB), reason='reference cpu.pb not present')


@needs_ref
def test_own_graphdef_reader_agrees_with_protobuf_library():
    from oracle.tf_graph import FrozenGraph
    from watsor_b200.graphdef import GraphDef
    ours, theirs = GraphDef(REF_PB), FrozenGraph(REF_PB)
    assert ours.order == theirs.order
    rng = np.random.default_rng(0)
    consts = [n for n in ours.order if ours.nodes[n].op == 'Const']
    for name in rng.choice(consts, 150, replace=False):
        a, b = ours.const(name), theirs.const(name)
        if a.dtype == object:
            assert [bytes(x) for x in a.ravel()] == [bytes(x) for x in np.asarray(b).ravel()]
            continue
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), name
    for name in ours.order[::97]:
        assert ours.nodes[name].op == theirs.nodes[name].op
        assert list(ours.nodes[name].input) == list(theirs.nodes[name].input)


@needs_ref
def test_compiled_program_matches_survey_appendix_a():
    m = compile_frozen_graph(REF_PB)
    assert (m.input_h, m.input_w, m.num_classes, m.num_anchors) == (300, 300, 3, 1917)
    assert abs(m.macs_per_frame / 1e6 - 1126.95) < 0.01
    assert [l.op for l in m.layers].count(OP_HEAD) == 6 and m.layers[0].op == OP_STEM
    assert (m.max_per_class, m.max_total) == (100, 100)
    assert np.float32(m.iou_thr) == np.float32(0.6) and np.float32(m.score_thr) == np.float32(0.3)
    assert [l.row_off for l in m.layers if l.op == OP_HEAD] == [0, 1083, 1683, 1833, 1887, 1911]
    # TF SAME padding is asymmetric: stem 300->150 pads nothing before, 75->38 pads one
    assert (m.layers[0].pad_t, m.layers[0].pad_l) == (0, 0)
    dw4 = [l for l in m.layers if l.in_h == 75 and l.out_h == 38][0]
    assert (dw4.pad_t, dw4.pad_l) == (1, 1)


@needs_ref
def test_anchors_fold_equals_oracle_and_formula():
    from oracle.ssd_graph import SsdGraphOracle
    m = compile_frozen_graph(REF_PB)
    assert np.array_equal(m.anchors, SsdGraphOracle(REF_PB).anchors)
    fm = [(19, 19), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)]
    assert np.array_equal(ssd_anchors(fm), m.anchors)


def check_arena(m):
    """No two simultaneously live tensors may overlap in the planned arena."""
    live = {}
    last = {}
    for i, l in enumerate(m.layers):
        for t in (l.src, l.src2):
            if t:
                last[t] = i
    for i, l in enumerate(m.layers[:-1]):       # liveness extension for fused pairs, as in plan_arena
        nxt = m.layers[i + 1]
        if l.op == 2 and nxt.op == 3 and nxt.src == l.dst and l.src:
            last[l.src] = max(last[l.src], i + 1)
    for i, l in enumerate(m.layers):
        if l.dst:
            size = l.out_h * l.out_w * l.out_c
            for t, (o, s) in live.items():
                assert l.out_off + size <= o or o + s <= l.out_off, (l.name, t)
            live[l.dst] = (l.out_off, size)
            assert l.out_off % 256 == 0 and l.out_off + size <= m.arena_elems
        for t in [t for t in live if last.get(t, -1) <= i and t != l.dst]:
            del live[t]
    # fused depthwise -> 1x1 pairs (csrc/kernels_fused.cu) read the depthwise input while writing the 1x1 output
    from watsor_b200.model import OP_DW, OP_PW
    for a, b in zip(m.layers, m.layers[1:]):
        if a.op == OP_DW and b.op == OP_PW and b.src == a.dst and a.src:
            a0, a1 = a.in_off, a.in_off + a.in_h * a.in_w * a.in_c
            b0, b1 = b.out_off, b.out_off + b.out_h * b.out_w * b.out_c
            assert a1 <= b0 or b1 <= a0, (a.name, b.name)


def test_arena_planner_and_blob_round_trip():
    m = synthetic_ssd_mobilenet_v1(num_classes=5, seed=3)
    check_arena(m)
    assert m.arena_elems < 3 * 150 * 150 * 64          # ping-pong, not one buffer per layer
    m2 = Model.from_blob(m.to_blob())
    assert len(m2.layers) == len(m.layers) and m2.arena_elems == m.arena_elems
    for a, b in zip(m.layers, m2.layers):
        assert (a.op, a.in_off, a.out_off, a.n_pad, a.row_off, a.name[:31]) == \
               (b.op, b.in_off, b.out_off, b.n_pad, b.row_off, b.name)
    for a, b in zip(m.tensors, m2.tensors):
        assert np.array_equal(a.ravel(), b.ravel())
    assert m2.num_anchors == 1917 and np.float32(m2.score_thr) == np.float32(1e-8)


def test_residual_add_keeps_both_inputs_alive():
    from watsor_b200.model import _Emitter
    m = Model()
    em = _Emitter(m)
    em.shape['image'] = (32, 32, 3)
    w = np.zeros((3, 3, 3, 16), np.float32)
    em.conv('stem', 'image', 'a', w, np.ones(16), np.zeros(16), 2, 1)
    em.conv('pw1', 'a', 'b', np.zeros((1, 1, 16, 16), np.float32), np.ones(16), np.zeros(16), 1, 0)
    em.add('add', 'a', 'b', 'c')
    em.conv('pw2', 'c', 'd', np.zeros((1, 1, 16, 16), np.float32), np.ones(16), np.zeros(16), 1, 0)
    m.plan_arena()
    check_arena(m)
    add = [l for l in m.layers if l.op == OP_ADD][0]
    assert add.in_off != add.in2_off


@needs_ref
def test_older_nms_exports_are_refused_not_miscompiled(tmp_path):
    """A graph whose post-processing is not the NonMaxSuppressionV5 topology (the 2018 model-zoo exports clip and
    filter per class before NonMaxSuppressionV2/V3) must raise instead of compiling with the wrong semantics."""
    from tensorboard.compat.proto import graph_pb2
    g = graph_pb2.GraphDef()
    with open(REF_PB, 'rb') as f:
        g.ParseFromString(f.read())
    for n in g.node:
        if n.op == 'NonMaxSuppressionV5':
            n.op = 'NonMaxSuppressionV3'
    p = tmp_path / 'old_export.pb'
    p.write_bytes(g.SerializeToString())
    with pytest.raises(NotImplementedError, match='NonMaxSuppressionV3'):
        compile_frozen_graph(str(p))
    # a per-class ClipToWindow_k in front of the NMS nodes is refused as well
    g.ParseFromString(open(REF_PB, 'rb').read())
    scope = next(n.name for n in g.node if n.op == 'NonMaxSuppressionV5').split('non_max_suppression')[0]
    extra = g.node.add()
    extra.name = scope + 'ClipToWindow_7/Minimum'
    extra.op = 'Identity'
    extra.input.append(scope + 'Minimum/x')
    p.write_bytes(g.SerializeToString())
    with pytest.raises(NotImplementedError, match='ClipToWindow_7'):
        compile_frozen_graph(str(p))
