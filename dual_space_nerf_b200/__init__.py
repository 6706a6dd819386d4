"""B200-native volume-rendering path for Dual-Space NeRF (drop-in for can_render.Renderer)."""
from .net import DualSpaceNeRF, synthetic_net  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torch.cuda / the .so
    if name == "Renderer":
        from .renderer import Renderer

        return Renderer
    raise AttributeError(name)
