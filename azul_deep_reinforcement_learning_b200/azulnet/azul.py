"""placeholder: filled in by the façade milestone."""


class IllegalMove(Exception):
    pass


class GameEnded(Exception):
    pass


class IllegalRule(Exception):
    pass


class Azul:
    pass
