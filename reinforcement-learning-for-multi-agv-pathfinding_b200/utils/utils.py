"""Action codes and small constant tables (reference: src/utils/utils.py:3-46)."""


class Direction:
    """UP 0, RIGHT 1, DOWN 2, LEFT 3, STOP 4 (utils.py:9-10); NN heads use the first
    action_num() == 4 of them (utils.py:12-15)."""

    NAMES = ("UP", "RIGHT", "DOWN", "LEFT", "STOP")
    DELTA = ((0, -1), (1, 0), (0, 1), (-1, 0), (0, 0))  # Explorer.py:127-129

    def __init__(self):
        self.value_str = dict(enumerate(self.NAMES))
        self.str_value = {n: i for i, n in enumerate(self.NAMES)}

    def action_num(self):
        return 4

    def action_str_value(self, action_str):
        return self.str_value[action_str]  # KeyError on a bad name, like the reference

    def action_value_str(self, action_value):
        return self.value_str[action_value]  # KeyError on a bad code (utils.py:21)


class ColorBox:
    """kept for API compatibility with the (out-of-scope) pygame renderer"""
    GRAY_Color = (192, 192, 192)
    BLACK_COLOR = (0, 0, 0)
    WHITE_COLOR = (255, 255, 255)
    RED_COLOR = (200, 30, 30)
    BLUE_COLOR = (30, 30, 200)
    PINK_COLOR = (255, 153, 204)
    GREEN_COLOR = (0, 255, 0)


class Working:
    """work-type tables (unused by the step logic, utils.py:35-46)"""

    def __init__(self):
        names = ["Start", "Halt", "Turning", "Lifting", "Downing", "Picking"]
        self.work_type_val_str = dict(enumerate(names))
        self.work_type_str_val = {n: i for i, n in enumerate(names)}
        self.work_time = dict(zip(names, (1, 1, 3, 6, 6, 16)))

    def time_return(self, work_type):
        return self.work_time[work_type]
