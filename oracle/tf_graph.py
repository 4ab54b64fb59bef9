"""GraphDef access for the oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Reads a frozen TensorFlow GraphDef with the protobuf classes that ship with
tensorboard (TensorFlow itself is not installed) and evaluates *constant*
sub-graphs with numpy.  The product has its own, independent wire-format reader
(watsor_b200/graphdef.py); tests cross-check the two.

Reference call site being restated: watsor/detection/tensorflow_cpu.py:50-62
(`od_graph_def.ParseFromString` + `tf.import_graph_def`).
"""
import numpy as np
from tensorboard.compat.proto import graph_pb2
from tensorboard.util import tensor_util


class FrozenGraph:
    def __init__(self, pb_path):
        g = graph_pb2.GraphDef()
        with open(pb_path, 'rb') as f:
            g.ParseFromString(f.read())
        self.nodes = {n.name: n for n in g.node}
        self.order = [n.name for n in g.node]
        self._cache = {}
        self.consumers = {}
        for n in g.node:
            for i in n.input:
                src = i.lstrip('^').split(':')[0]
                self.consumers.setdefault(src, []).append(n.name)

    # ------------------------------------------------------------------ helpers
    def node(self, name):
        return self.nodes[name]

    def inputs(self, name):
        """Data inputs (control dependencies dropped) as (node, output_index)."""
        out = []
        for i in self.nodes[name].input:
            if i.startswith('^'):
                continue
            parts = i.split(':')
            out.append((parts[0], int(parts[1]) if len(parts) > 1 else 0))
        return out

    def const(self, name):
        """Value of a Const node, following Identity (`.../read`) chains."""
        n = self.nodes[name]
        while n.op == 'Identity':
            n = self.nodes[self.inputs(n.name)[0][0]]
        assert n.op == 'Const', (name, n.op)
        return tensor_util.make_ndarray(n.attr['value'].tensor)

    def ops(self, op):
        return [name for name in self.order if self.nodes[name].op == op]

    # --------------------------------------------------- constant-folding eval
    def eval(self, ref):
        """Evaluate a constant sub-graph with numpy (fp32 stays fp32).

        Implements exactly the op set that appears under the
        `MultipleGridAnchorGenerator/` and `Concatenate/` scopes of a TF
        Object-Detection SSD graph.  Element-wise fp32 IEEE ops (+,-,*,/,sqrt) are
        correctly rounded in numpy as in TF's Eigen kernels, so the result is the
        graph's own value, bit for bit.
        """
        parts = ref.split(':')
        name, idx = parts[0], int(parts[1]) if len(parts) > 1 else 0
        key = (name, idx)
        if key in self._cache:
            return self._cache[key]
        n = self.nodes[name]
        args = [self.eval('%s:%d' % i) for i in self.inputs(name)]
        op = n.op
        if op == 'Const':
            v = tensor_util.make_ndarray(n.attr['value'].tensor)
        elif op == 'Identity':
            v = args[0]
        elif op in ('AddV2', 'Add'):
            v = args[0] + args[1]
        elif op == 'Sub':
            v = args[0] - args[1]
        elif op == 'Mul':
            v = args[0] * args[1]
        elif op == 'RealDiv':
            v = args[0] / args[1]
        elif op == 'Sqrt':
            v = np.sqrt(args[0])
        elif op == 'Minimum':
            v = np.minimum(args[0], args[1])
        elif op == 'Maximum':
            v = np.maximum(args[0], args[1])
        elif op == 'Cast':
            dst = tensor_util.dtypes.as_dtype(n.attr['DstT'].type).as_numpy_dtype
            v = np.asarray(args[0]).astype(dst)
        elif op == 'Range':
            v = np.arange(args[0], args[1], args[2], dtype=np.asarray(args[0]).dtype)
        elif op == 'Reshape':
            v = np.reshape(args[0], [int(x) for x in np.asarray(args[1]).ravel()])
        elif op == 'ExpandDims':
            v = np.expand_dims(args[0], int(args[1]))
        elif op == 'Tile':
            v = np.tile(args[0], [int(x) for x in np.asarray(args[1]).ravel()])
        elif op == 'Pack':
            v = np.stack(args, axis=int(n.attr['axis'].i))
        elif op == 'ConcatV2':
            v = np.concatenate([np.asarray(a) for a in args[:-1]], axis=int(args[-1]))
        elif op == 'Fill':
            v = np.full([int(x) for x in np.asarray(args[0]).ravel()], args[1],
                        dtype=np.asarray(args[1]).dtype)
        elif op == 'Slice':
            begin = [int(x) for x in np.asarray(args[1]).ravel()]
            size = [int(x) for x in np.asarray(args[2]).ravel()]
            sl = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
            v = np.asarray(args[0])[sl]
        elif op == 'StridedSlice':
            v = self._strided_slice(n, args)
        elif op == 'Shape':
            v = np.array(np.asarray(args[0]).shape, dtype=np.int32)
        else:
            raise NotImplementedError('oracle const-eval: op %s (%s)' % (op, name))
        v = np.asarray(v)
        self._cache[key] = v
        return v

    @staticmethod
    def _strided_slice(n, args):
        x = np.asarray(args[0])
        begin, end, strides = (np.asarray(a).ravel() for a in args[1:4])
        bm, em = n.attr['begin_mask'].i, n.attr['end_mask'].i
        sm = n.attr['shrink_axis_mask'].i
        assert n.attr['ellipsis_mask'].i == 0 and n.attr['new_axis_mask'].i == 0
        idx = []
        for d in range(len(begin)):
            if sm & (1 << d):
                idx.append(int(begin[d]))
                continue
            b = None if bm & (1 << d) else int(begin[d])
            e = None if em & (1 << d) else int(end[d])
            idx.append(slice(b, e, int(strides[d])))
        return x[tuple(idx)]
