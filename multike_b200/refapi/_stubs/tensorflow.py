"""Stand-in for `import tensorflow as tf` in the reference's HOST-side modules when TensorFlow 1.x is
not installed (it has no wheel for current Pythons).  The only uses outside the modules this package
replaces are utils.load_session (utils.py:25-28: ConfigProto + Session) and
tf.global_variables_initializer().run(session=) in the two drivers, which this package re-authors.
Put this directory on sys.path AFTER multike_b200/refapi and only if `import tensorflow` fails."""


class _Options:
    allow_growth = False


class ConfigProto:
    def __init__(self, *args, **kwargs):
        self.gpu_options = _Options()


class Session:
    def __init__(self, config=None, **kwargs):
        self.config = config

    def close(self):
        pass


class _Initializer:
    def run(self, session=None):
        pass


def global_variables_initializer():
    return _Initializer()
