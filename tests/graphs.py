"""Graph builders shared by tests, smoke() and bench.py. `d` is a DSL namespace (oracle or
exprgrad_b200.frontend), `L` the matching layer library; both expose the same names. `ct` is the
compile target of every target ("gpu" for the product; the CPU baseline uses "threads" = the reference's
row-split thread pool, exprgrad/parser.nim:800-810)."""
import numpy as np


def matmul(d, L, ct="gpu"):
    """benchmarks/matmul/matmul_gpu.nim:28-36"""
    c = d.Fun(); y, x, it = d.Iter("y"), d.Iter("x"), d.Iter("it")
    c[y, x] += d.input("a")[y, it] * d.input("b")[it, x]
    return [c.target("c", ct)]


def dense_net(d, L, sizes=(784, 512, 512, 10), rate=0.01, ct="gpu"):
    """BASELINE config 3: dense+relu stack, softmax + crossEntropy, gradientDescent
    (exprgrad/layers/dnn.nim:19-27, 90-94; base.nim:37-38, 66-67)."""
    x = d.input("x", [-1, sizes[0]]); y = d.input("y", [-1, sizes[-1]])
    h = x
    for i in range(len(sizes) - 2):
        h = L.relu(L.dense(h, sizes[i], sizes[i + 1]))
    p = L.softmax(L.dense(h, sizes[-2], sizes[-1]))
    loss = L.cross_entropy(p, y)
    return [p.target("predict", ct), loss.target("loss", ct),
            loss.backprop(L.gradient_descent(rate)).target("train", ct)]


def dense_inputs(batch, sizes=(784, 512, 512, 10), seed=0):
    """SURVEY.md 8(d) C3 inputs: x ~ U(0,1), labels one-hot(randint), params U(-0.1, 0.1)."""
    x = np.random.default_rng(seed).uniform(0, 1, (batch, sizes[0])).astype(np.float32)
    lab = np.random.default_rng(seed + 1).integers(0, sizes[-1], batch)
    y = np.zeros((batch, sizes[-1]), np.float32)
    y[np.arange(batch), lab] = 1
    rng = np.random.default_rng(seed + 2)
    params = []
    for i in range(len(sizes) - 1):
        params.append(rng.uniform(-0.1, 0.1, (sizes[i], sizes[i + 1])).astype(np.float32))
        params.append(rng.uniform(-0.1, 0.1, (sizes[i + 1],)).astype(np.float32))
    return x, y, params


def xor_net(d, L, rate=0.1, ct="gpu"):
    """examples/xor/xor.nim:20-28 (BASELINE config 1)"""
    net = L.sigmoid(L.dense(L.leaky_relu(L.dense(d.input("x"), 2, 4)), 4, 1)).target("predict", ct)
    loss = L.mse(net, d.input("y")).target("loss", ct)
    return [loss.backprop(L.gradient_descent(rate)).target("train", ct)]


def conv2_net(d, L, ct="gpu", filters=(4, 3, 3, 3)):
    """benchmarks/conv2 style: NHWC valid convolution forward + backward to filters and images."""
    img = d.input("img"); w = d.param(list(filters), name="filters")
    out = L.conv2(img, w)
    loss = d.Fun(); it = d.Iter("it")
    loss[0] += d.sq(out.raw[it])
    return [out.target("conv", ct), loss.target("loss", ct),
            loss.backwards().grad(w).target("dw", ct), loss.backwards().grad(img).target("dimg", ct)]


def fashion_net(d, L, rate=0.01, ct="gpu"):
    """examples/fashion_mnist/fashion_mnist.nim:40-57 shape: conv -> leakyRelu -> maxpool -> dense, adam."""
    x = d.input("x", [-1, 12, 12, 1]); y = d.input("y", [-1, 10])
    h = L.maxpool2(L.leaky_relu(L.conv2_layer(x, 1, 3, 3, 4)))          # [N,5,5,4]
    h = h.reshape([-1, 100])
    p = L.softmax(L.dense(h, 100, 10))
    loss = L.cross_entropy(p, y)
    return [p.target("predict", ct), loss.target("loss", ct), loss.backprop(L.adam(rate)).target("train", ct)]


ALL = {"matmul": matmul, "dense_net": dense_net, "xor_net": xor_net, "conv2_net": conv2_net, "fashion_net": fashion_net}
