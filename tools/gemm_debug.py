"""Accuracy of one 1x1-conv GEMM in every precision mode against a float64 reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from watsor_b200.engine import Engine
from watsor_b200.model import Model, _Emitter


def tiny_model(K, N, hw, seed=0):
    rng = np.random.default_rng(seed)
    m = Model(name='gemm-test', input_h=hw, input_w=hw, num_classes=1, num_anchors=1)
    em = _Emitter(m)
    em.shape['image'] = (hw, hw, 3)
    w0 = rng.standard_normal((1, 1, 3, K)).astype(np.float32)
    em.conv('stem', 'image', 'a', w0, np.ones(K, np.float32), np.zeros(K, np.float32), 1, 0)
    w1 = (rng.standard_normal((1, 1, K, N)) / np.sqrt(K)).astype(np.float32)
    em.conv('pw', 'a', 'b', w1, np.ones(N, np.float32), np.zeros(N, np.float32), 1, 0)
    m.anchors_tensor = m.add_tensor(np.zeros((1, 4), np.float32))
    m.plan_arena()
    return m, w1.reshape(K, N)


for (K, N, hw, n) in [(512, 512, 19, 4), (32, 64, 32, 2), (1024, 1024, 10, 8), (256, 48, 3, 1), (64, 128, 20, 3)]:
    m, w1 = tiny_model(K, N, hw)
    rng = np.random.default_rng(1)
    pre = rng.standard_normal((n, hw, hw, 3)).astype(np.float32)
    out = {}
    for p in (0, 3, 2, 1):
        with Engine(m.to_blob(), device=0, max_batch=n, precision=p) as e:
            _, _, a = e.backbone(pre, stop_layer=0, layer_shape=(hw, hw, K))
            _, _, y = e.backbone(pre, stop_layer=1, layer_shape=(hw, hw, N))
            out[p] = (a, y)
    a32 = out[0][0].reshape(-1, K)
    for p in (0, 3, 2, 1):
        a, y = out[p]
        ref = a.reshape(-1, K).astype(np.float64) @ w1.astype(np.float64)
        d = y.reshape(-1, N).astype(np.float64) - ref
        print('K%5d N%5d M%6d prec %d: max|err| %.3e  mean err %+.3e  rms %.3e  (|ref| rms %.3f)  A same as fp32: %s' % (
            K, N, ref.shape[0], p, np.abs(d).max(), d.mean(), np.sqrt((d ** 2).mean()), np.sqrt((ref ** 2).mean()),
            np.array_equal(a.reshape(-1, K), a32)), flush=True)
        if p in (2, 3):
            sgn = np.sign(ref)
            print('        signed-toward-zero bias: mean(err*sign(ref)) = %+.3e' % (d * sgn).mean())
