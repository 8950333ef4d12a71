"""theano.tensor.shared_randomstreams — only constructed, never used, by convRBM.py:156."""


class RandomStreams(object):
    def __init__(self, seed=None):
        self.seed = seed
