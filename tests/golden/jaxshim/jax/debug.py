def print(fmt, *a, **k):
    __builtins__["print"](fmt.format(*a, **k)) if isinstance(__builtins__, dict) else None


def callback(f, *a, **k):
    f(*a, **k)
