class EzPickle:
    """gym.utils.EzPickle: remembers constructor arguments (GraspingEnv.py:43)."""

    def __init__(self, *args, **kwargs):
        self._ezpickle_args = args
        self._ezpickle_kwargs = kwargs
