"""pygame stand-in: import-time only (render_mode=None in every oracle use). TEST INFRASTRUCTURE."""


class _Missing:
    def __getattr__(self, name):
        raise RuntimeError("pygame is not available: rendering is out of scope for the oracle")


image = transform = display = event = mouse = font = draw = surfarray = _Missing()
SRCALPHA = 0


def init():
    raise RuntimeError("pygame is not available")


def quit():
    return None
