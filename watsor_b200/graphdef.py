"""Dependency-free reader for frozen TensorFlow GraphDef files (`cpu.pb`,
`frozen_inference_graph.pb`) -- the model files watsor selects by name in
watsor/detection/tensorflow_cpu.py:50-53 -- plus a small constant folder.

The reference hands the file to TensorFlow (`od_graph_def.ParseFromString`,
tensorflow_cpu.py:55-60).  Here the protobuf wire format is decoded directly
(GraphDef / NodeDef / AttrValue / TensorProto field numbers from the public
tensorflow/core/framework/*.proto definitions) so that the model compiler
(watsor_b200/model.py) needs neither TensorFlow nor tensorboard.
"""
import struct

import numpy as np

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 4, 7, 9, 10
_NP = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_UINT8: np.uint8,
       DT_INT64: np.int64, DT_BOOL: np.bool_}


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message; length-delimited
    values are returned as memoryview slices (zero copy)."""
    pos = 0
    end = len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield fn, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v, wt):
    if wt == 0:
        return [_signed(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


class Tensor:
    __slots__ = ('dtype', 'shape', 'content', 'float_val', 'int_val', 'int64_val', 'bool_val',
                 'double_val', 'string_val')

    def numpy(self):
        if self.dtype == DT_STRING:
            return np.array(self.string_val, dtype=object).reshape(self.shape)
        npdt = _NP[self.dtype]
        n = int(np.prod(self.shape)) if self.shape else 1
        if self.content is not None and len(self.content):
            arr = np.frombuffer(self.content, dtype=npdt)
        else:
            vals = {DT_FLOAT: self.float_val, DT_DOUBLE: self.double_val, DT_INT32: self.int_val,
                    DT_UINT8: self.int_val, DT_INT64: self.int64_val,
                    DT_BOOL: self.bool_val}[self.dtype]
            arr = np.asarray(vals, dtype=npdt)
            if arr.size == 0:
                arr = np.zeros(n, dtype=npdt)
            elif arr.size < n:       # TensorProto: the last value repeats
                arr = np.concatenate([arr, np.full(n - arr.size, arr[-1], dtype=npdt)])
        return arr.reshape(self.shape).copy()


def _parse_tensor(buf):
    t = Tensor()
    t.dtype, t.shape, t.content = 0, [], None
    t.float_val, t.int_val, t.int64_val, t.bool_val, t.double_val, t.string_val = [], [], [], [], [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            t.dtype = v
        elif fn == 2:                                   # TensorShapeProto
            for f2, _, v2 in _fields(v):
                if f2 == 2:                             # Dim
                    size = 0
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            size = _signed(v3)
                    t.shape.append(size)
        elif fn == 4:
            t.content = bytes(v)
        elif fn == 5:
            t.float_val += ([struct.unpack('<f', v)[0]] if wt == 5 else
                            list(np.frombuffer(bytes(v), '<f4')))
        elif fn == 6:
            t.double_val += ([struct.unpack('<d', v)[0]] if wt == 1 else
                             list(np.frombuffer(bytes(v), '<f8')))
        elif fn == 7:
            t.int_val += _packed_varints(v, wt)
        elif fn == 8:
            t.string_val.append(bytes(v))
        elif fn == 10:
            t.int64_val += _packed_varints(v, wt)
        elif fn == 11:
            t.bool_val += [bool(x) for x in _packed_varints(v, wt)]
    return t


def _parse_attr(buf):
    """AttrValue -> python value (only the kinds an inference graph uses)."""
    for fn, wt, v in _fields(buf):
        if fn == 2:
            return bytes(v)
        if fn == 3:
            return _signed(v)
        if fn == 4:
            return struct.unpack('<f', v)[0]
        if fn == 5:
            return bool(v)
        if fn == 6:
            return ('type', v)
        if fn == 8:
            return _parse_tensor(v)
        if fn == 1:                                     # ListValue
            out = []
            for f2, w2, v2 in _fields(v):
                if f2 == 3:
                    out += _packed_varints(v2, w2)
                elif f2 == 2:
                    out.append(bytes(v2))
                elif f2 == 4:
                    out += ([struct.unpack('<f', v2)[0]] if w2 == 5 else
                            list(np.frombuffer(bytes(v2), '<f4')))
            return out
    return None


class Node:
    __slots__ = ('name', 'op', 'input', 'attr')

    def data_inputs(self):
        out = []
        for i in self.input:
            if i.startswith('^'):
                continue
            p = i.split(':')
            out.append((p[0], int(p[1]) if len(p) > 1 else 0))
        return out


class GraphDef:
    def __init__(self, path):
        with open(path, 'rb') as f:
            buf = memoryview(f.read())
        self.nodes = {}
        self.order = []
        for fn, _, v in _fields(buf):
            if fn != 1:
                continue
            n = Node()
            n.name, n.op, n.input, n.attr = '', '', [], {}
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    n.name = bytes(v2).decode()
                elif f2 == 2:
                    n.op = bytes(v2).decode()
                elif f2 == 3:
                    n.input.append(bytes(v2).decode())
                elif f2 == 5:
                    key, val = None, None
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            key = bytes(v3).decode()
                        elif f3 == 2:
                            val = v3
                    n.attr[key] = val               # parsed lazily (weights are big)
            self.nodes[n.name] = n
            self.order.append(n.name)
        self.consumers = {}
        for name in self.order:
            for src, _ in self.nodes[name].data_inputs():
                self.consumers.setdefault(src, []).append(name)
        self._fold = {}

    def attr(self, name, key, default=None):
        raw = self.nodes[name].attr.get(key)
        if raw is None:
            return default
        if isinstance(raw, memoryview):
            raw = _parse_attr(raw)
            self.nodes[name].attr[key] = raw
        return raw

    def ops(self, op):
        return [n for n in self.order if self.nodes[n].op == op]

    def const(self, name):
        n = self.nodes[name]
        while n.op == 'Identity':
            n = self.nodes[n.data_inputs()[0][0]]
        if n.op != 'Const':
            raise ValueError('%s is not constant (%s)' % (name, n.op))
        return self.attr(n.name, 'value').numpy()

    # ------------------------------------------------------------- const folding
    def fold(self, name):
        """Evaluate a constant sub-graph (the SSD anchor generator) with numpy.

        Element-wise float32 IEEE arithmetic is correctly rounded in numpy, so the
        folded anchors are the values TensorFlow's own constant folding yields.
        """
        if name in self._fold:
            return self._fold[name]
        n = self.nodes[name]
        a = [self.fold(src) for src, _ in n.data_inputs()]
        op = n.op
        if op == 'Const':
            v = self.attr(name, 'value').numpy()
        elif op == 'Identity':
            v = a[0]
        elif op in ('Add', 'AddV2'):
            v = a[0] + a[1]
        elif op == 'Sub':
            v = a[0] - a[1]
        elif op == 'Mul':
            v = a[0] * a[1]
        elif op == 'RealDiv':
            v = a[0] / a[1]
        elif op == 'Sqrt':
            v = np.sqrt(a[0])
        elif op == 'Minimum':
            v = np.minimum(a[0], a[1])
        elif op == 'Maximum':
            v = np.maximum(a[0], a[1])
        elif op == 'Cast':
            v = np.asarray(a[0]).astype(_NP[self.attr(name, 'DstT')[1]])
        elif op == 'Range':
            v = np.arange(a[0], a[1], a[2], dtype=np.asarray(a[0]).dtype)
        elif op == 'Reshape':
            v = np.reshape(a[0], [int(x) for x in np.ravel(a[1])])
        elif op == 'ExpandDims':
            v = np.expand_dims(a[0], int(a[1]))
        elif op == 'Tile':
            v = np.tile(a[0], [int(x) for x in np.ravel(a[1])])
        elif op == 'Pack':
            v = np.stack(a, axis=int(self.attr(name, 'axis', 0)))
        elif op == 'ConcatV2':
            v = np.concatenate([np.asarray(x) for x in a[:-1]], axis=int(a[-1]))
        elif op == 'Fill':
            v = np.full([int(x) for x in np.ravel(a[0])], a[1], dtype=np.asarray(a[1]).dtype)
        elif op == 'Slice':
            begin = [int(x) for x in np.ravel(a[1])]
            size = [int(x) for x in np.ravel(a[2])]
            v = np.asarray(a[0])[tuple(slice(b, None if s == -1 else b + s)
                                       for b, s in zip(begin, size))]
        elif op == 'StridedSlice':
            begin, end, strides = (np.ravel(x) for x in a[1:4])
            bm = self.attr(name, 'begin_mask', 0)
            em = self.attr(name, 'end_mask', 0)
            sm = self.attr(name, 'shrink_axis_mask', 0)
            idx = []
            for d in range(len(begin)):
                if sm >> d & 1:
                    idx.append(int(begin[d]))
                else:
                    idx.append(slice(None if bm >> d & 1 else int(begin[d]),
                                     None if em >> d & 1 else int(end[d]), int(strides[d])))
            v = np.asarray(a[0])[tuple(idx)]
        elif op == 'Shape':
            v = np.array(np.shape(a[0]), dtype=np.int32)
        else:
            raise NotImplementedError('constant folding: op %s (%s)' % (op, name))
        v = np.asarray(v)
        self._fold[name] = v
        return v
