"""Data formats either side of the CUDA hot path.

Two formats exist:

* the **packed** device format: ``uint32 state[W(P)][G]`` structure-of-arrays, W(P) = 7 + 5·P
  words per game (68 / 88 / 108 bytes for P = 2 / 3 / 4 players) -- what every kernel reads and
  writes in HBM (see DESIGN.md "HBM layout" for the bit assignment), and
* the **unpacked** interchange record: ``int32 rec[G][U(P)]``, U(P) = 48 + 58·P, which mirrors
  the attribute set of the reference ``Azul`` object one-to-one (``azulnet/azul.py:18-61``) plus
  the fields the reference keeps but ``export_JSON`` drops (box/lid, ``end_of_game``, the
  statistics arrays, ``azul.py:31-33,51-52,58``).  ``azb_import_state`` / ``azb_export_state``
  (kernel K7) convert between the two on the device; tests, the golden vectors and the
  ``azulnet`` façade speak the unpacked record.

Everything here is plain offsets -- no game logic lives in this module.
"""
from dataclasses import dataclass

N_DISPLAYS = 5           # azul.py:19 -- always five factories, whatever the player count
N_COLORS = 5
N_ACTIONS = 180          # 6 sources x 5 colours x 6 destinations, game_runner.py:102-103
MASK_WORDS = 6           # word p holds the 30 (source, colour) bits of destination p

TILE_POOL_RANDOM = 0     # azul.py:45-47
TILE_POOL_LID = 1        # azul.py:48-52
FIRST_PLAYER_RANDOM = 0  # azul.py:36-37 ; 1..P = fixed seat (azul.py:38-39, default 1 azul.py:43)

# per-game status bits written by the kernels (replace the reference's exceptions)
STATUS_ILLEGAL = 1       # IllegalMove, azul.py:301-302 -- state left untouched
STATUS_ENDED = 2         # GameEnded,  azul.py:298-299 -- state left untouched
STATUS_STUCK = 4         # no legal action and round not over (reference crashes)
STATUS_BAG_EMPTY = 8     # Lid pool: box and lid both empty during refill (azul.py:86 TODO)
STATUS_BAD_IMPORT = 16   # azb_import_state: record not representable in the packed format


def state_words(players: int) -> int:
    """uint32 words per game in the packed SoA format."""
    return 7 + 5 * players


def state_bytes(players: int) -> int:
    return 4 * state_words(players)


def algorithmic_bytes_per_step(players: int) -> int:
    """packed state read + write, 1-byte action in, 24-byte next mask out."""
    return 2 * state_bytes(players) + 1 + 4 * MASK_WORDS


@dataclass(frozen=True)
class UnpackedLayout:
    """int32 offsets of the unpacked record for ``players`` seats."""
    players: int

    @property
    def displays(self):            # [5][5] count of colour c on factory i   (azul.py:19)
        return 0

    @property
    def center(self):              # [6] colour counts + first-player token  (azul.py:20)
        return 25

    @property
    def pattern_lines(self):       # [P][5][5] count per (row, colour)        (azul.py:21-22)
        return 31

    @property
    def walls(self):               # [P][5][5] bool per (row, COLOUR)         (azul.py:23-24)
        return 31 + 25 * self.players

    @property
    def floors(self):              # [P] 0..7                                  (azul.py:25)
        return 31 + 50 * self.players

    @property
    def score(self):               # [P]                                       (azul.py:26)
        return 31 + 51 * self.players

    @property
    def scalars(self):             # current_player, next_first_player, players, end_of_game, turn_counter
        return 31 + 52 * self.players

    @property
    def current_player(self):
        return self.scalars + 0

    @property
    def next_first_player(self):
        return self.scalars + 1

    @property
    def n_players(self):
        return self.scalars + 2

    @property
    def end_of_game(self):
        return self.scalars + 3

    @property
    def turn_counter(self):
        return self.scalars + 4

    @property
    def box(self):                 # [5] (Lid pool only, else 0)               (azul.py:51)
        return self.scalars + 5

    @property
    def lid(self):                 # [5]                                       (azul.py:52)
        return self.scalars + 10

    @property
    def first_player_stats(self):  # [P]                                       (azul.py:31)
        return self.scalars + 15

    @property
    def floor_penalty(self):       # [P] cumulative, <= 0 like the reference   (azul.py:32,208)
        return self.first_player_stats + self.players

    @property
    def max_combo(self):           # [P]                                       (azul.py:33,264)
        return self.first_player_stats + 2 * self.players

    @property
    def completed_lines(self):     # [P][3] 0 row, 1 colour, 2 column          (azul.py:58)
        return self.first_player_stats + 3 * self.players

    @property
    def total_steps(self):         # env steps this slot has executed (RNG position; not in the reference)
        return self.first_player_stats + 6 * self.players

    @property
    def status(self):              # sticky STATUS_* bits (not in the reference)
        return self.total_steps + 1

    @property
    def size(self):
        return self.status + 1


def unpacked_size(players: int) -> int:
    return 48 + 58 * players


for _p in (2, 3, 4):
    assert UnpackedLayout(_p).size == unpacked_size(_p)
