"""Global flags (reference: src/utils/settings.py:4-18).  Read when a Scene is built."""


class SuperParas:
    Explorer_Always_Loaded = False  # every AGV counts as loaded (cannot pass under shelves)
    Explorer_Always_Empty = False   # every AGV counts as empty
    FPS = 300                       # renderer only
    Sparse_Reward = False           # the reference's name for the SHAPED reward (Explorer.py:206-249)
