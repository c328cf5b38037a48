"""Seeded random `++=` graphs for differential tests (native passes vs the oracle's restatement of passes.nim on CPU;
device kernels vs the oracle's loop nests on the GPU). The same seed builds the same graph with either DSL module
(`oracle` or `exprgrad_b200.frontend`), like tests/graphs.py.

A graph is a chain of 2-4 kernels over inputs a, b : [R, C], v : [C] and parameters w : [C, 5], u : [C]:
raw maps, 2-D maps with a row broadcast, shifted reads (conv1-like, test_model.nim:91-97), row / column
reductions, a contraction, then a scalar loss; targets: the chain's result, the loss, d loss / d a, and an SGD
step on the parameters the chain used. Expressions are random trees over every differentiable opcode of
ir.nim:51-76 whose adjoint `derive` knows (passes.nim:383-517), kept inside their domains (sqrt / ln / pow / div see
e*e + 1) so that the numeric comparison is meaningful; comparisons only look at input values (exact on both sides),
never at computed ones, so that a last-ulp difference cannot flip a branch."""
import random

COLS = 6   # C: w is [C, 5], u is [C]
KOUT = 5


def _expr(d, rng, leaves, cmp_leaves, depth):
    """random expression over `leaves` (Expr objects of this kernel)"""
    if depth <= 0 or rng.random() < 0.15:
        if rng.random() < 0.2:
            return d.lift(rng.choice([0.5, 2.0, -1.5, 3.0, 0.25]))
        return rng.choice(leaves)
    sub = lambda: _expr(d, rng, leaves, cmp_leaves, depth - 1)
    op = rng.choice(["add", "sub", "mul", "div", "neg", "sin", "cos", "exp", "sqrt", "ln", "pow", "select", "max", "min",
                     "sq", "scale"])
    if op == "add": return sub() + sub()
    if op == "sub": return sub() - sub()
    if op == "mul": return sub() * sub()
    if op == "div":
        den = sub()
        return sub() / (den * den + 1.0)
    if op == "neg": return -sub()
    if op == "sin": return d.sin(sub())
    if op == "cos": return d.cos(sub())
    if op == "exp": return d.exp(d.sin(sub()))
    if op == "sqrt":
        e = sub()
        return d.sqrt(e * e + 1.0)
    if op == "ln":
        e = sub()
        return d.ln(e * e + 1.0)
    if op == "pow":
        e = sub()
        return d.pow_(e * e + 1.0, rng.choice([1.5, 2.0, 0.5]))
    if op == "select":
        x, y = rng.choice(cmp_leaves), rng.choice(cmp_leaves)
        cond = (x < y) if rng.random() < 0.5 else (x <= d.lift(rng.choice([0.0, 0.25, -0.5])))
        return d.select(cond, sub(), sub())
    if op == "max": return d.max_(rng.choice(cmp_leaves), rng.choice(cmp_leaves)) * sub()
    if op == "min": return d.min_(rng.choice(cmp_leaves), rng.choice(cmp_leaves)) + sub()
    if op == "sq": return d.sq(sub())
    return sub() * rng.choice([0.5, 2.0, -1.0])


def _through(rng, e, leaf):
    """keep the chain differentiable: the kernel's value depends on `leaf` whatever the random tree picked"""
    return e * leaf if rng.random() < 0.5 else e + leaf


def random_net(d, L, seed, ct="gpu", depth=3):
    """-> (graphs, description). Targets: out, loss, da, and train when a parameter was used."""
    rng = random.Random(seed)
    a = d.input("a", [-1, COLS]); b = d.input("b", [-1, COLS]); v = d.input("v", [COLS])
    w = d.param([COLS, KOUT], name="w"); u = d.param([COLS], name="u")
    cur, kind = a, "RC"      # RC: [R, C] like a and b; RK: [R, 5]; RS: [R, C - 2]; R: [R]; C: [C]
    used_param = False
    steps = []
    for _ in range(rng.randint(2, 4)):
        r = d.Fun()
        if kind == "RC":
            op = rng.choice(["raw", "row", "shift", "rowsum", "colsum", "contract", "bias"])
        elif kind in ("RK", "RS"):
            op = rng.choice(["raw1", "rowsum", "colsum"])
        else:
            op = "raw1"
        steps.append(op)
        if op == "raw":          # r{i} += E(cur{i}, b{i}, a{i})
            it = d.Iter("it")
            leaves = [cur.raw[it], b.raw[it]]
            r.raw[it] += _through(rng, _expr(d, rng, leaves, [a.raw[it], b.raw[it]], depth), leaves[0])
            r.copy_shape(cur)
        elif op == "raw1":       # r{i} += E(cur{i})
            it = d.Iter("it")
            x = cur.raw[it]
            r.raw[it] += _through(rng, _expr(d, rng, [x], [d.lift(0.5), d.lift(-0.25)], depth - 1), x)
            r.copy_shape(cur)
        elif op == "row":        # r[y, x] += E(cur[y, x], v[x], u[x])
            y, x = d.Iter("y"), d.Iter("x")
            leaves = [cur[y, x], v[x], u[x]]
            r[y, x] += _through(rng, _expr(d, rng, leaves, [a[y, x], v[x]], depth), leaves[0])
            used_param = True
        elif op == "bias":       # two kernels writing one tensor (dnn.nim:19-24)
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += cur[y, x] * v[x]
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += u[x]
            used_param = True
        elif op == "shift":      # r[y, x] += E(cur[y, x + 1], cur[y, x + 2]): x stands alone only in the write, so its
            # bounds follow r's inferred shape [R, C - 2] (a read cur[y, x] would bound x by C and run x + 1 off the row,
            # in the reference too: passes.nim:1119-1180 takes the first access an iterator indexes alone)
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += _through(rng, _expr(d, rng, [cur[y, x + 1], cur[y, x + 2]], [a[y, x + 1], b[y, x + 2]], depth - 1), cur[y, x + 1])
            kind = "RS"
        elif op == "rowsum":     # r[y] += E(cur[y, x])
            y, x = d.Iter("y"), d.Iter("x")
            r[y] += _through(rng, _expr(d, rng, [cur[y, x]], [d.lift(0.5), d.lift(-0.25)], depth - 1), cur[y, x])
            kind = "R"
        elif op == "colsum":     # r[x] += E(cur[y, x])
            y, x = d.Iter("y"), d.Iter("x")
            r[x] += _through(rng, _expr(d, rng, [cur[y, x]], [d.lift(0.5), d.lift(-0.25)], depth - 1), cur[y, x])
            kind = "C"
        else:                    # contraction r[y, k] += cur[y, x] * w[x, k]
            y, k, x = d.Iter("y"), d.Iter("k"), d.Iter("x")
            r[y, k] += cur[y, x] * w[x, k]
            kind = "RK"
            used_param = True
        cur = r
    loss = d.Fun(); it = d.Iter("it")
    x = cur.raw[it]
    loss[0] += _expr(d, rng, [x], [d.lift(0.5), d.lift(-0.25)], 2) * x
    graphs = [cur.target("out", ct), loss.target("loss", ct), loss.backwards().grad(a).target("da", ct)]
    if used_param:
        graphs.append(loss.backprop(L.gradient_descent(0.05)).target("train", ct))
    return graphs, " > ".join(steps)


def random_inputs(np, seed, rows=7):
    rng = np.random.default_rng(1000 + seed)
    f = lambda *s: rng.uniform(-1, 1, s).astype(np.float32)
    return {"a": f(rows, COLS), "b": f(rows, COLS), "v": f(COLS)}


def random_params(np, seed):
    rng = np.random.default_rng(2000 + seed)
    return {"w": rng.uniform(-1, 1, (COLS, KOUT)).astype(np.float32), "u": rng.uniform(-1, 1, (COLS,)).astype(np.float32)}


def random_cnn(d, L, seed, ct="gpu"):
    """Random composition of the reference's layer library (layers/dnn.nim:19-100, base.nim:37-67) - CPU-tier passes
    parity only: conv2 / activations / max- and average pooling (customGrad, strided and divided indices) / upsample
    (`withShape`) / dropout (`TensorRandom`) / reshape generator / dense / softmax, a loss and an optimizer (adam adds
    caches and `epoch()`). -> (graphs, description, input shapes)."""
    rng = random.Random(10_000 + seed)
    side = rng.choice([8, 12, 16])
    chans = rng.choice([1, 3])
    x = d.input("x", [-1, side, side, chans])
    cur, h, c = x, side, chans
    steps = []
    for _ in range(rng.randint(1, 4)):
        op = rng.choice(["conv", "act", "maxpool", "avgpool", "upsample", "dropout"])
        if op == "conv" and h >= 5:
            k = rng.choice([3, 3, 5]) if h >= 7 else 3
            f = rng.choice([2, 4])
            cur = L.conv2_layer(cur, c, k, k, f); h, c = h - k + 1, f
        elif op == "act":
            cur = getattr(L, rng.choice(["relu", "leaky_relu", "sigmoid", "tanh"]))(cur)
        elif op == "maxpool" and h >= 4 and h % 2 == 0:
            cur = L.maxpool2(cur); h //= 2
        elif op == "avgpool" and h >= 4 and h % 2 == 0:
            cur = L.avgpool2(cur); h //= 2
        elif op == "upsample" and h <= 8:
            cur = L.upsample2(cur); h *= 2
        elif op == "dropout":
            cur = L.dropout(cur, 0.25)
        else:
            continue
        steps.append(op)
    flat = h * h * c
    cur = cur.reshape([-1, flat])
    hidden = rng.choice([0, 8])
    if hidden:
        cur = L.relu(L.dense(cur, flat, hidden)); flat = hidden
    outs = rng.choice([1, 4])
    logits = L.dense(cur, flat, outs)
    y = d.input("y", [-1, outs])
    head = rng.choice(["softmax", "sigmoid", "none"]) if outs > 1 else rng.choice(["sigmoid", "none"])
    if head == "softmax":
        p = L.softmax(logits); loss = L.cross_entropy(p, y)
    elif head == "sigmoid":
        p = L.sigmoid(logits); loss = rng.choice([L.binary_cross_entropy, L.mse])(p, y)
    else:
        p = logits; loss = L.mse(p, y)
    opt = L.adam(0.01) if rng.random() < 0.5 else L.gradient_descent(0.05)
    steps += [f"reshape[{flat}]", head, "adam" if "adam" in repr(opt) or opt.__qualname__.startswith("adam") else "sgd"]
    graphs = [p.target("predict", ct), loss.target("loss", ct), loss.backprop(opt).target("train", ct)]
    return graphs, " > ".join(steps), {"x": [rng.choice([1, 5]), side, side, chans], "y": None, "outs": outs}


def random_index_net(d, L, seed, ct="gpu"):
    """Kernels whose INDICES do the work - explicit loop bounds, strided / divided / modular / wrapped accesses, iterator
    and shape values inside expressions, array literals, scatter writes, `withShape` (ir.nim:51-76 IndexDiv / Mod / Wrap /
    ToScalar / Shape / Len / Array*; test_model.nim:91-154, 233-263 hold one small case of each) - chained two or three deep,
    with a loss and its gradient. Inputs: a : [R, 12], v : [12]. -> (graphs, description)."""
    rng = random.Random(20_000 + seed)
    a = d.input("a", [-1, 12]); v = d.input("v", [12])
    cur, kind = a, "2d"          # "2d": [R, C] with C a multiple of 6 unless noted; "1d": [n]
    steps = []
    for _ in range(rng.randint(1, 3)):
        r = d.Fun()
        if kind == "2d":
            op = rng.choice(["stride", "updiv", "modrow", "iterval", "array", "rowstencil", "colmean", "flatten_wrap"])
        else:
            op = rng.choice(["stencil", "wrap", "scatter", "lenmean"])
        steps.append(op)
        if op == "stride":           # r[y, x] += sum_k c_k * cur[y, s x + k]: strided reads, shape [R, C / s]
            s = rng.choice([2, 3]); y, x = d.Iter("y"), d.Iter("x")
            e = None
            for k in range(s):
                t = cur[y, x * s + k] * rng.choice([0.5, 1.0, -2.0])
                e = t if e is None else e + t
            r[y, x] += e
            kind = "2d_odd"
        elif op == "updiv":          # r[y, x] += cur[y, x div 2] * c: divided index, explicit shape (dnn.nim:81-88)
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += cur[y, x // 2] * rng.choice([1.0, 0.25])
            r.with_shape(cur.shape[0], cur.shape[1] * 2)
        elif op == "modrow":         # r[y, x] += cur[y, x] * v[x mod 3]
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += cur[y, x] * v[x % 3]
        elif op == "iterval":        # iterator and shape values inside the expression
            y, x = d.Iter("y"), d.Iter("x")
            r[y, x] += cur[y, x] * d.to_scalar(x + 1) / d.to_scalar(cur.shape[1])
        elif op == "array":          # array literal indexed by an expression of the iterator
            y, x = d.Iter("y"), d.Iter("x")
            arr = d.lift([rng.choice([0.5, 1.5, -1.0]) for _ in range(3)])
            r[y, x] += cur[y, x] * arr[x % 3] + d.to_scalar(d.array_len(arr))
        elif op == "rowstencil":     # explicit bounds on the column loop: r[y, x - 1] += ...
            y = d.Iter("y"); x = d.Iter("x", 1, cur.shape[1] - 1)
            r[y, x - 1] += (cur[y, x - 1] + cur[y, x] * 2.0 + cur[y, x + 1]) / 4.0
            kind = "2d_odd"
        elif op == "colmean":        # r[x] += cur[y, x] / rows
            y, x = d.Iter("y"), d.Iter("x")
            r[x] += cur[y, x] / d.to_scalar(cur.shape[0])
            kind = "1d"
        elif op == "flatten_wrap":   # r{i} += cur{wrap(i + k, len)}: a circular shift of the flattened tensor
            it = d.Iter("it")
            r.raw[it] += cur.raw[d.wrap(it + rng.choice([1, 5]), cur.len())] * 0.5
            r.copy_shape(cur)
        elif op == "stencil":        # 1-D blur with explicit bounds (test_model.nim:109-117)
            x = d.Iter("x", 1, cur.shape[0] - 1)
            r[x - 1] += (cur[x - 1] + cur[x] + cur[x + 1]) / 3.0
        elif op == "wrap":           # circular shift
            x = d.Iter("x")
            r[x] += cur[d.wrap(x + rng.choice([1, 2, 7]), cur.shape[0])] - cur[x] * 0.5
            r.copy_shape(cur)
        elif op == "scatter":        # r[x div 2] += cur[x]: the write index is not an iterator
            x = d.Iter("x")
            r[x // 2] += cur[x] * rng.choice([1.0, -0.5])
        else:                        # lenmean: r[0] += cur[x] / len
            x = d.Iter("x")
            r[0] += cur[x] / d.to_scalar(cur.len())
            kind = "scalar"
        cur = r
        if kind in ("2d_odd", "scalar"):
            break
    loss = d.Fun(); it = d.Iter("it")
    loss[0] += d.sq(cur.raw[it]) * 0.5 + cur.raw[it]
    return [cur.target("out", ct), loss.target("loss", ct), loss.backwards().grad(a).target("da", ct)], " > ".join(steps)
