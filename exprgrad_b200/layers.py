"""Layer library of exprgrad (exprgrad/layers/base.nim:19-67, exprgrad/layers/dnn.nim:19-100) expressed
with the exprgrad_b200 front-end, so that the BASELINE configurations (dense + relu + softmax +
crossEntropy + gradientDescent, conv2, ...) can be built without a Nim toolchain. Iterator creation
order inside each kernel follows the reference expressions (it fixes loop and accumulation order).
Only graph construction happens here; execution is libegb200.so.
"""
from .frontend import (Fun, Iter, cache, epoch, exp, ln, max_, min_, param, pow_, rand, select, sin as sin_, sq, sqrt,
                 to_scalar, grad_arg, KernelBuilder, lift)


def _named(f, name):
    f.name = name
    return f


def add(a, b):
    r = Fun(); it = Iter("it"); r.raw[it] += a.raw[it] + b.raw[it]; return _named(r, "+")
def sub(a, b):
    r = Fun(); it = Iter("it"); r.raw[it] += a.raw[it] - b.raw[it]; return _named(r, "-")
def minimum(a, b):
    r = Fun(); it = Iter("it"); r.raw[it] += min_(a.raw[it], b.raw[it]); return _named(r, "min")
def maximum(a, b):
    r = Fun(); it = Iter("it"); r.raw[it] += max_(a.raw[it], b.raw[it]); return _named(r, "max")
def scale(a, factor):
    r = Fun(); it = Iter("it"); r.raw[it] += a.raw[it] * float(factor); return _named(r, "*")
def divide(a, factor):
    r = Fun(); it = Iter("it"); r.raw[it] += a.raw[it] / float(factor); return _named(r, "/")


def matmul(a, b):  # base.nim:27-28
    r = Fun(); y, x, it = Iter("y"), Iter("x"), Iter("it")
    r[y, x] += a[y, it] * b[it, x]
    return _named(r, "matmul")


def transpose(m):  # base.nim:32-33
    r = Fun(); y, x = Iter("y"), Iter("x")
    r[y, x] += m[x, y]
    return _named(r, "transpose")


def gradient_descent(rate=0.01):  # base.nim:37-38
    def opt(p, g):
        it = Iter("it")
        p.raw[it] += -g.raw[it] * float(rate)
    return opt


def adam(eta=0.01, beta1=0.9, beta2=0.999, eps=1e-8):  # base.nim:40-53
    def opt(p, g):
        m, v = cache(p, "adam.m"), cache(p, "adam.v")
        it = Iter("it")
        m.raw[it] += m.raw[it] * (lift(beta1) - 1.0) + (lift(1.0) - beta1) * g.raw[it]
        it = Iter("it")
        v.raw[it] += v.raw[it] * (lift(beta2) - 1.0) + (lift(1.0) - beta2) * sq(g.raw[it])
        it = Iter("it")
        m_hat = m.raw[it] / (1.0 - pow_(beta1, to_scalar(epoch())))
        v_hat = v.raw[it] / (1.0 - pow_(beta2, to_scalar(epoch())))
        p.raw[it] += -lift(eta) * m_hat / (sqrt(v_hat) + eps)
    return opt


def mse(a, b):  # base.nim:57-58
    r = Fun(); it = Iter("it")
    r[0] += sq(a.raw[it] - b.raw[it]) / to_scalar(a.shape[0])
    return _named(r, "mse")


def binary_cross_entropy(pred, labels):  # base.nim:60-64
    r = Fun(); it = Iter("it")
    r[0] += -(labels.raw[it] * ln(pred.raw[it]) + (1.0 - labels.raw[it]) * ln(1.0 - pred.raw[it])) / to_scalar(pred.shape[0])
    return _named(r, "binaryCrossEntropy")


def cross_entropy(pred, labels):  # base.nim:66-67
    r = Fun(); it = Iter("it")
    r[0] += -(labels.raw[it] * ln(pred.raw[it])) / to_scalar(pred.shape[0])
    return _named(r, "crossEntropy")


def dense(values, inp, outp, has_bias=True):  # dnn.nim:19-24
    w = param([inp, outp], name="weights")
    r = Fun(); y, x, it = Iter("y"), Iter("x"), Iter("it")
    r[y, x] += values[y, it] * w[it, x]
    if has_bias:
        b = param([outp], name="bias")
        y, x = Iter("y"), Iter("x")
        r[y, x] += b[x]
    return _named(r, "dense")


def relu(inp):  # dnn.nim:26-27
    r = Fun(); it = Iter("it")
    r.raw[it] += select(inp.raw[it] >= 0.0, inp.raw[it], 0.0)
    return _named(r, "relu")


def leaky_relu(inp, leak=0.01):  # dnn.nim:29-30
    r = Fun(); it = Iter("it")
    r.raw[it] += select(inp.raw[it] >= 0.0, 1.0, float(leak)) * inp.raw[it]
    return _named(r, "leakyRelu")


def sigmoid(inp):  # dnn.nim:32-33
    r = Fun(); it = Iter("it")
    r.raw[it] += 1.0 / (1.0 + exp(-inp.raw[it]))
    return _named(r, "sigmoid")


def tanh(inp):  # dnn.nim:35-40
    r = Fun(); it = Iter("it")
    a = exp(inp.raw[it]); b = exp(-inp.raw[it])
    r.raw[it] += (a - b) / (a + b)
    return _named(r, "tanh")


def sin(inp):  # dnn.nim:42-43
    r = Fun(); it = Iter("it")
    r.raw[it] += sin_(inp.raw[it])
    return _named(r, "sin")


def conv2(images, filters):  # dnn.nim:45-49 (NHWC, filters [filter, dy, dx, chan], valid)
    r = Fun()
    image, y, x, f, dx, dy, chan = (Iter(n) for n in ("image", "y", "x", "filter", "dx", "dy", "chan"))
    r[image, y, x, f] += images[image, y + dy, x + dx, chan] * filters[f, dy, dx, chan]
    return _named(r, "conv2")


def conv2_layer(images, chans, w, h, filters):  # dnn.nim:51-53
    return conv2(images, param([filters, h, w, chans], name="filters"))


def maxpool2(images):  # dnn.nim:58-71 (customGrad)
    r = Fun()
    image, y, x, chan = (Iter(n) for n in ("image", "y", "x", "chan"))
    value = max_(max_(images[image, y * 2, x * 2, chan], images[image, y * 2 + 1, x * 2, chan]),
                 max_(images[image, y * 2, x * 2 + 1, chan], images[image, y * 2 + 1, x * 2 + 1, chan]))
    gi, gy, gx, gc = (Iter(n) for n in ("image", "y", "x", "chan"))
    gvalue = select(images[gi, gy, gx, gc].eq(r[gi, gy // 2, gx // 2, gc]),
                    grad_arg(r)[gi, gy // 2, gx // 2, gc], 0.0)
    gk = KernelBuilder(grad_arg(images), [lift(i, "index") for i in (gi, gy, gx, gc)], lift(gvalue), False)
    r.add_kernel([lift(i, "index") for i in (image, y, x, chan)], value, False, custom_grad=[gk])
    r.lock()
    return _named(r, "maxpool2")


def avgpool2(images):  # dnn.nim:73-79
    r = Fun()
    image, y, x, chan = (Iter(n) for n in ("image", "y", "x", "chan"))
    r[image, y, x, chan] += (images[image, y * 2, x * 2, chan] + images[image, y * 2 + 1, x * 2, chan] +
                             images[image, y * 2, x * 2 + 1, chan] + images[image, y * 2 + 1, x * 2 + 1, chan]) / 4.0
    return _named(r, "avgpool2")


def upsample2(images):  # dnn.nim:81-88
    r = Fun()
    image, y, x, chan = (Iter(n) for n in ("image", "y", "x", "chan"))
    r[image, y, x, chan] += images[image, y // 2, x // 2, chan]
    r.with_shape(images.shape[0], images.shape[1] * 2, images.shape[2] * 2, images.shape[3])
    return _named(r, "upsample2")


def softmax(inp):  # dnn.nim:90-94 (no max-subtraction)
    sums = Fun(); y, x = Iter("y"), Iter("x")
    sums[y] += exp(inp[y, x])
    sums.name = "softmax.sums"
    r = Fun(); y, x = Iter("y"), Iter("x")
    r[y, x] += exp(inp[y, x]) / sums[y]
    return _named(r, "softmax")


def dropout(inp, prob):  # dnn.nim:96-100
    rnd = rand(inp, (0.0, 1.0)); rnd.name = "dropout.rand"
    r = Fun(); it = Iter("it")
    r.raw[it] += select(lift(float(prob)) <= rnd.raw[it], inp.raw[it] / (1.0 - float(prob)), 0.0)
    r.copy_shape(inp)
    return _named(r, "dropout")
