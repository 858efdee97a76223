class CenteredNorm:
    def __init__(self, *args, **kwargs):
        pass
