"""xframe_b200 -- B200-native (sm_100a) drop-in for the MTIP phasing hot path of `xframe fxs reconstruct`."""
__all__ = ['Plan', 'HIO', 'ER']


def __getattr__(name):
    if name in __all__:
        from . import plan
        return getattr(plan, name)
    raise AttributeError(name)
