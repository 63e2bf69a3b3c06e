class Rotation:  # only named in an import on the reference's path; never called by the golden generator
    def __init__(self, *a, **k):
        raise NotImplementedError
