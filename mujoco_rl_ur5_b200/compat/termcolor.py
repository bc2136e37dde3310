"""Minimal stand-in for `termcolor.colored` (reference: GraspingEnv.py:20,106-150; MujocoController.py:9)."""
_COLORS = dict(grey=30, red=31, green=32, yellow=33, blue=34, magenta=35, cyan=36, white=37)
_ATTRS = dict(bold=1, dark=2, underline=4, blink=5, reverse=7, concealed=8)


def colored(text, color=None, on_color=None, attrs=None):
    codes = []
    if color in _COLORS:
        codes.append(str(_COLORS[color]))
    if on_color and on_color.startswith("on_") and on_color[3:] in _COLORS:
        codes.append(str(_COLORS[on_color[3:]] + 10))
    for a in attrs or []:
        if a in _ATTRS:
            codes.append(str(_ATTRS[a]))
    return f"\033[{';'.join(codes)}m{text}\033[0m" if codes else str(text)


def cprint(text, color=None, on_color=None, attrs=None, **kw):
    print(colored(text, color, on_color, attrs), **kw)
