/*
 * azul_oracle.c -- CPU restatement of the reference Azul rules.   TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA hot path, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or call
 * it.  It restates /root/reference/azulnet/azul.py and the mask / random-agent parts of
 * /root/reference/azulnet/game_runner.py in plain C on the reference's own UNPACKED arrays
 * (deliberately not the packed device format, so that it is an independent statement of the
 * rules).  Every function cites the reference lines it follows.
 *
 * Parity pin: tests/test_oracle_golden.py replays the tests/golden npz files (recorded from the
 * unmodified reference by oracle/record_golden.py) through this file: every post-step record,
 * every 180-bit legal mask, the reference's board fixtures and the known-answer scenarios of
 * the reference's tests/test_azul.py must match exactly.
 *
 * The parts that have no counterpart in the reference -- the Philox4x32-10 draw schedule, the
 * integer random-agent sampler, auto-reset and the rollout counters -- are the SPEC shared with
 * the CUDA kernels (DESIGN.md "RNG schedule"); they are restated here independently so the two
 * implementations can be compared bit for bit on seeded rollouts.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAXP 4
#define MAXF 9

/* Number of factory displays.  The reference always uses 5 (azul.py:19, tests/test_azul.py:13-15; its azul.py:72 is a
 * TODO); the opt-in variant of the engine ("factory count by player count", SURVEY §8f rank 4: 5 / 7 / 9 displays for
 * 2 / 3 / 4 players as in the board game) is checked against this same file with g_F set through ao_set_factories().
 * With g_F = 5 every function below is the reference-pinned default.  Process-global: test infrastructure only. */
static int g_F = 5;
void ao_set_factories(int f) { g_F = (f == 7 || f == 9) ? f : 5; }
int ao_get_factories(void) { return g_F; }

typedef struct {
    int32_t displays[MAXF][5];       /* azul.py:19 */
    int32_t center[6];               /* azul.py:20 */
    int32_t pattern_lines[MAXP][5][5];/* azul.py:21-22 */
    int32_t walls[MAXP][5][5];       /* azul.py:23-24, indexed [row][COLOUR] */
    int32_t floors[MAXP];            /* azul.py:25 */
    int32_t score[MAXP];             /* azul.py:26 */
    int32_t current_player;          /* azul.py:27, 1-based, 0 = unset */
    int32_t next_first_player;       /* azul.py:37-43 */
    int32_t players;                 /* azul.py:28 */
    int32_t end_of_game;             /* azul.py:29 */
    int32_t turn_counter;            /* azul.py:30 */
    int32_t box[5], lid[5];          /* azul.py:51-52 (Lid pool only) */
    int32_t first_player_stats[MAXP];/* azul.py:31 */
    int32_t floor_penalty[MAXP];     /* azul.py:32 (<= 0) */
    int32_t max_combo[MAXP];         /* azul.py:33 */
    int32_t completed_lines[MAXP][3];/* azul.py:58: 0 row, 1 colour, 2 column */
    uint32_t total_steps;            /* spec: env steps executed by this slot (RNG position) */
    int32_t status;                  /* spec: sticky status bits */
} ao_game;

enum { ST_ILLEGAL = 1, ST_ENDED = 2, ST_STUCK = 4, ST_BAG_EMPTY = 8 };
enum { POOL_RANDOM = 0, POOL_LID = 1 };
enum { PURPOSE_ACTION = 0, PURPOSE_REFILL = 1, PURPOSE_FIRST = 2, PURPOSE_RESET_REFILL = 3 };

/* ------------------------------------------------------------------ record <-> struct ---- */

int ao_record_size(int players) { return 48 + 58 * players + 5 * (g_F - 5); }

static void from_record(ao_game *g, const int32_t *r, int P)
{
    memset(g, 0, sizeof(*g));
    const int32_t *p = r;
    memcpy(g->displays, p, 5 * g_F * 4); p += 5 * g_F;
    memcpy(g->center, p, 6 * 4); p += 6;
    for (int i = 0; i < P; i++) { memcpy(g->pattern_lines[i], p, 25 * 4); p += 25; }
    for (int i = 0; i < P; i++) { memcpy(g->walls[i], p, 25 * 4); p += 25; }
    memcpy(g->floors, p, P * 4); p += P;
    memcpy(g->score, p, P * 4); p += P;
    g->current_player = *p++; g->next_first_player = *p++; g->players = *p++;
    g->end_of_game = *p++; g->turn_counter = *p++;
    memcpy(g->box, p, 20); p += 5;
    memcpy(g->lid, p, 20); p += 5;
    memcpy(g->first_player_stats, p, P * 4); p += P;
    memcpy(g->floor_penalty, p, P * 4); p += P;
    memcpy(g->max_combo, p, P * 4); p += P;
    for (int i = 0; i < P; i++) { memcpy(g->completed_lines[i], p, 12); p += 3; }
    g->total_steps = (uint32_t)*p++;
    g->status = *p++;
    g->players = P;
}

static void to_record(const ao_game *g, int32_t *r, int P)
{
    int32_t *p = r;
    memcpy(p, g->displays, 5 * g_F * 4); p += 5 * g_F;
    memcpy(p, g->center, 6 * 4); p += 6;
    for (int i = 0; i < P; i++) { memcpy(p, g->pattern_lines[i], 25 * 4); p += 25; }
    for (int i = 0; i < P; i++) { memcpy(p, g->walls[i], 25 * 4); p += 25; }
    memcpy(p, g->floors, P * 4); p += P;
    memcpy(p, g->score, P * 4); p += P;
    *p++ = g->current_player; *p++ = g->next_first_player; *p++ = P;
    *p++ = g->end_of_game; *p++ = g->turn_counter;
    memcpy(p, g->box, 20); p += 5;
    memcpy(p, g->lid, 20); p += 5;
    memcpy(p, g->first_player_stats, P * 4); p += P;
    memcpy(p, g->floor_penalty, P * 4); p += P;
    memcpy(p, g->max_combo, P * 4); p += P;
    for (int i = 0; i < P; i++) { memcpy(p, g->completed_lines[i], 12); p += 3; }
    *p++ = (int32_t)g->total_steps;
    *p++ = g->status;
}

/* ------------------------------------------------------------------ Philox4x32-10 -------- */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); Random123 constants. */

void ao_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void draw_words(uint64_t seed, uint32_t gid, uint32_t n, uint32_t purpose, uint32_t j, uint32_t out[4])
{
    uint32_t ctr[4] = { gid, n, purpose, j };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    ao_philox4x32_10(ctr, key, out);
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ------------------------------------------------------------------ rules ---------------- */

/* Python negative indexing: arrays indexed with current_player-1 wrap to the last seat when
 * current_player == 0 (fresh object before new_round, azul.py:27; tests/test_game_runner.py:92) */
static inline int seat(const ao_game *g) { return g->current_player > 0 ? g->current_player - 1 : g->players - 1; }

/* azul.py:18-61 */
static void init_game(ao_game *g, int P, int pool, int first_player)
{
    uint32_t T = g->total_steps;
    memset(g, 0, sizeof(*g));
    g->players = P;
    g->next_first_player = first_player;              /* azul.py:35-43 */
    if (pool == POOL_LID) for (int c = 0; c < 5; c++) g->box[c] = 20;   /* azul.py:51-52 */
    g->total_steps = T;
}

/* azul.py:64-89 with the 20 colours supplied (draws[i*4+j], azul.py:74-75 order).
 * Lid pool bookkeeping follows azul.py:81-89: refill check before EVERY draw. */
static void new_round_injected(ao_game *g, int pool, const int8_t *draws)
{
    g->current_player = g->next_first_player;                          /* :66 */
    if (g->next_first_player >= 1) g->first_player_stats[g->next_first_player - 1] += 1; /* :67 */
    else g->first_player_stats[g->players - 1] += 1;                   /* python [-1] */
    g->turn_counter += 1;                                              /* :68 */
    g->next_first_player = 0;                                          /* :69 */
    for (int c = 0; c < 5; c++) g->center[c] = 0;                      /* :71 */
    g->center[5] = 1;
    memset(g->displays, 0, sizeof(g->displays));                       /* :73 */
    for (int i = 0; i < g_F; i++)
        for (int j = 0; j < 4; j++) {
            int c = draws[i * 4 + j];
            if (c < 0) continue;                                       /* spec: slot left empty */
            if (pool == POOL_LID) {
                int sum = 0;
                for (int k = 0; k < 5; k++) sum += g->box[k];
                if (sum == 0) {                                        /* :81-83 */
                    for (int k = 0; k < 5; k++) { g->box[k] = g->lid[k]; g->lid[k] = 0; }
                }
                g->displays[i][c] += 1;                                /* :88 */
                g->box[c] -= 1;                                        /* :89 */
            } else {
                g->displays[i][c] += 1;                                /* :78 */
            }
        }
}

/* SPEC draw schedule (DESIGN.md "RNG schedule"): turn Philox words into the 20 colours.
 * Random pool: display i uses word R[i]; colour = mulhi(x,5), x *= 5, four times.
 * Lid pool: display i uses words R[2i], R[2i+1], two draws each: r = mulhi(x,total), x *= total,
 * colour = first c with box[0]+..+box[c] > r (a without-replacement draw proportional to the box
 * counts, azul.py:85-89), lid poured into the box when the box is empty (azul.py:81-83). */
static void new_round_philox(ao_game *g, int pool, uint64_t seed, uint32_t gid, uint32_t purpose)
{
    uint32_t R[20];
    int calls = pool == POOL_LID ? (2 * g_F + 3) / 4 : (g_F + 3) / 4;       /* 3 / 2 calls for the reference's 5 displays */
    for (int j = 0; j < calls; j++) draw_words(seed, gid, g->total_steps, purpose, (uint32_t)j, R + 4 * j);
    int8_t draws[4 * MAXF];
    if (pool == POOL_RANDOM) {
        for (int i = 0; i < g_F; i++) {
            uint32_t x = R[i];
            for (int j = 0; j < 4; j++) { draws[i * 4 + j] = (int8_t)mulhi32(x, 5); x *= 5u; }
        }
        new_round_injected(g, pool, draws);
        return;
    }
    /* Lid: the colour depends on the evolving box, so interleave with the bookkeeping */
    int32_t box[5], lid[5];
    memcpy(box, g->box, sizeof(box)); memcpy(lid, g->lid, sizeof(lid));
    for (int i = 0; i < g_F; i++) {
        uint32_t x = 0;
        for (int j = 0; j < 4; j++) {
            if ((j & 1) == 0) x = R[2 * i + (j >> 1)];
            int sum = box[0] + box[1] + box[2] + box[3] + box[4];
            if (sum == 0) {
                for (int k = 0; k < 5; k++) { box[k] = lid[k]; lid[k] = 0; }
                sum = box[0] + box[1] + box[2] + box[3] + box[4];
            }
            if (sum == 0) { draws[i * 4 + j] = -1; g->status |= ST_BAG_EMPTY; continue; }
            uint32_t r = mulhi32(x, (uint32_t)sum);
            x *= (uint32_t)sum;
            int c = 0, cum = box[0];
            while ((uint32_t)cum <= r) { c++; cum += box[c]; }
            draws[i * 4 + j] = (int8_t)c;
            box[c] -= 1;
        }
    }
    new_round_injected(g, pool, draws);
}

/* azul.py:119-123 */
static void add_to_floor(ao_game *g, int n)
{
    int s = seat(g);
    if (g->floors[s] + n < 7) g->floors[s] += n; else g->floors[s] = 7;
}

/* azul.py:118-161 */
static void move(ao_game *g, int pool, int display, int color, int pattern)
{
    int s = seat(g), nr_tiles;
    if (display != 0) {                                                /* :125-133 */
        nr_tiles = g->displays[display - 1][color];
        g->displays[display - 1][color] = 0;
        for (int c = 0; c < 5; c++) g->center[c] += g->displays[display - 1][c];
        for (int c = 0; c < 5; c++) g->displays[display - 1][c] = 0;
    } else {                                                           /* :134-143 */
        nr_tiles = g->center[color];
        g->center[color] = 0;
        if (g->center[5] == 1) {
            g->center[5] = 0;
            g->next_first_player = g->current_player;
            add_to_floor(g, 1);
        }
    }
    if (pattern != 0) {                                                /* :145-157 */
        int overflow = pattern - g->pattern_lines[s][pattern - 1][color] - nr_tiles;
        if (overflow >= 0) {
            g->pattern_lines[s][pattern - 1][color] += nr_tiles;
        } else {
            g->pattern_lines[s][pattern - 1][color] = pattern;
            add_to_floor(g, -overflow);
            if (pool == POOL_LID) g->lid[color] += -overflow;
        }
    } else {                                                           /* :158-161 */
        add_to_floor(g, nr_tiles);
        if (pool == POOL_LID) g->lid[color] += nr_tiles;
    }
}

/* azul.py:162-176 */
static int is_legal_move(const ao_game *g, int display, int color, int pattern)
{
    int s = seat(g);
    if (display > 0) { if (g->displays[display - 1][color] < 1) return 0; }
    else             { if (g->center[color] < 1) return 0; }
    if (pattern != 0) {
        for (int c = 0; c < 5; c++)
            if (c != color && g->pattern_lines[s][pattern - 1][c] != 0) return 0;     /* :172-173 */
        if (g->walls[s][pattern - 1][color]) return 0;                                 /* :174-175 */
    }
    return 1;
}

/* azul.py:177-181 */
static void next_player(ao_game *g)
{
    if (g->current_player < g->players) g->current_player += 1; else g->current_player = 1;
}

/* azul.py:182-183 -- all six centre slots count, token included */
static int is_end_of_round(const ao_game *g)
{
    for (int i = 0; i < g_F; i++) for (int c = 0; c < 5; c++) if (g->displays[i][c]) return 0;
    for (int c = 0; c < 6; c++) if (g->center[c]) return 0;
    return 1;
}

/* azul.py:184-191 */
static int is_end_of_game(const ao_game *g)
{
    for (int p = 0; p < g->players; p++)
        for (int i = 0; i < 5; i++) {
            int n = 0;
            for (int c = 0; c < 5; c++) n += g->walls[p][i][c] != 0;
            if (n == 5) return 1;
        }
    return 0;
}

static inline int to_wall_position(int color, int pattern) { return ((color + pattern) % 5 + 5) % 5; }   /* :194-196 */
static inline int from_wall_position(int color, int pattern) { return ((color - pattern) % 5 + 5) % 5; } /* :197-199 */

/* azul.py:200-210 */
static int count_floor(ao_game *g, int player)
{
    int f = g->floors[player], count;
    if (f <= 2) count = -f;
    else if (f <= 5) count = -2 - (f - 2) * 2;
    else count = -8 - (f - 5) * 3;
    g->floor_penalty[player] += count;
    g->floors[player] = 0;
    return count;
}

/* azul.py:211-290 */
static int count_wall(ao_game *g, int pool, int player)
{
    int count = 0;
    for (int pattern = 0; pattern < 5; pattern++)
        for (int color = 0; color < 5; color++) {
            if (g->pattern_lines[player][pattern][color] != pattern + 1) continue;     /* :216 */
            g->pattern_lines[player][pattern][color] = 0;                               /* :218 */
            g->walls[player][pattern][color] = 1;                                       /* :219 */
            if (pool == POOL_LID) g->lid[color] += pattern;                             /* :220-222 */
            int pos_count = 0, bonus_count = 0, only_row = 1, only_col = 1;
            int w = to_wall_position(color, pattern);
            for (int i = w + 1; i < 5; i++) {                                           /* :230-236 */
                if (g->walls[player][pattern][from_wall_position(i, pattern)]) { pos_count++; only_row = 0; }
                else break;
            }
            for (int i = w - 1; i >= 0; i--) {                                          /* :237-242 */
                if (g->walls[player][pattern][from_wall_position(i, pattern)]) { pos_count++; only_row = 0; }
                else break;
            }
            for (int j = pattern + 1; j < 5; j++) {                                     /* :244-250 */
                if (g->walls[player][j][to_wall_position(color, pattern - j)]) { pos_count++; only_col = 0; }
                else break;
            }
            for (int j = pattern - 1; j >= 0; j--) {                                    /* :251-257 */
                if (g->walls[player][j][to_wall_position(color, pattern - j)]) { pos_count++; only_col = 0; }
                else break;
            }
            if (only_row && only_col) pos_count = 1;                                    /* :258-263 */
            else if (!(only_row || only_col)) pos_count += 2;
            else pos_count += 1;
            if (pos_count > g->max_combo[player]) g->max_combo[player] = pos_count;     /* :264 */
            for (int i = 0; i < 5; i++) {                                               /* :266-272 */
                if (g->walls[player][pattern][i]) { if (i == 4) { bonus_count += 2; g->completed_lines[player][0]++; } }
                else break;
            }
            for (int j = 0; j < 5; j++) {                                               /* :274-280 */
                if (g->walls[player][j][color]) { if (j == 4) { bonus_count += 10; g->completed_lines[player][1]++; } }
                else break;
            }
            for (int k = 0; k < 5; k++) {                                               /* :282-288 */
                if (g->walls[player][k][from_wall_position(w, k)]) { if (k == 4) { bonus_count += 7; g->completed_lines[player][2]++; } }
                else break;
            }
            count += pos_count + bonus_count;                                           /* :289 */
        }
    return count;
}

/* azul.py:192,291-295 */
static void count_score(ao_game *g, int pool)
{
    for (int p = 0; p < g->players; p++) {
        int f = count_floor(g, p);
        int w = count_wall(g, pool, p);
        g->score[p] += f + w;
        if (g->score[p] < 0) g->score[p] = 0;
    }
}

/* game_runner.py:113-117 with index = d + S*c + 5*S*p, S = displays + 1 sources (game_runner.py:102-103: S = 6) ->
 * six words of 5*S bits, word p bit (d + S*c) */
static void legal_mask64(const ao_game *g, uint64_t mask[6])
{
    const int S = g_F + 1;
    for (int p = 0; p < 6; p++) mask[p] = 0;
    for (int i = 0; i < 30 * S; i++) {
        int d = i % S, c = (i / S) % 5, p = i / (5 * S);                                /* :107-111 */
        if (is_legal_move(g, d, c, p)) mask[p] |= 1ull << (d + S * c);
    }
}
static void legal_mask(const ao_game *g, uint32_t mask[6])
{
    uint64_t m[6];
    legal_mask64(g, m);
    for (int p = 0; p < 6; p++) mask[p] = (uint32_t)m[p];
}

/* azul.py:296-313.  draws == NULL -> Philox schedule.  Returns 0, or -1 IllegalMove / -2 GameEnded
 * with the state untouched (azul.py:298-302). */
static int step(ao_game *g, int pool, int d, int c, int p, const int8_t *draws, uint64_t seed, uint32_t gid)
{
    if (g->end_of_game) return -2;
    if (d < 0 || d > g_F || c < 0 || c > 4 || p < 0 || p > 5) return -1;
    if (!is_legal_move(g, d, c, p)) return -1;
    move(g, pool, d, c, p);
    g->total_steps += 1;
    if (is_end_of_round(g)) {
        count_score(g, pool);
        if (is_end_of_game(g)) g->end_of_game = 1;
        else if (draws) new_round_injected(g, pool, draws);
        else new_round_philox(g, pool, seed, gid, PURPOSE_REFILL);
    } else {
        next_player(g);
    }
    return 0;
}

/* SPEC: fresh game in a slot (Azul(...) + new_round(), game_runner.py:79-80), Philox draws.
 * first_rule 0 -> random.choice(1..P) (azul.py:37) as 1 + mulhi(word0, P); else the fixed seat. */
static void reset_philox(ao_game *g, int P, int pool, int first_rule, uint64_t seed, uint32_t gid)
{
    int first = first_rule;
    if (first_rule == 0) {
        uint32_t w[4];
        draw_words(seed, gid, g->total_steps, PURPOSE_FIRST, 0, w);
        first = 1 + (int)mulhi32(w[0], (uint32_t)P);
    }
    init_game(g, P, pool, first);
    new_round_philox(g, pool, seed, gid, PURPOSE_RESET_REFILL);
}

/* SPEC of the integer random agent: game_runner.py:87-97 gives weight 1.0 to every legal action
 * except the 30 straight-to-floor ones (p = 0) which get 0.01 -> integer weights 100 : 1.
 * r = mulhi(word, total) picks a point in the cumulative weight; heavy actions (words 1..5 in
 * ascending action index) come first, then the floor actions of word 0. */
static int random_action64(const uint64_t mask[6], uint32_t word)
{
    const int S = g_F + 1;
    int n_hi = 0;
    for (int p = 1; p < 6; p++) n_hi += __builtin_popcountll(mask[p]);
    int n_lo = __builtin_popcountll(mask[0]);
    uint32_t total = 100u * (uint32_t)n_hi + (uint32_t)n_lo;
    if (total == 0) return -1;
    uint32_t r = mulhi32(word, total);
    int k, p0, p1;
    if (r < 100u * (uint32_t)n_hi) { k = (int)(r / 100u); p0 = 1; p1 = 6; }
    else { k = (int)(r - 100u * (uint32_t)n_hi); p0 = 0; p1 = 1; }
    for (int p = p0; p < p1; p++)
        for (int b = 0; b < 5 * S; b++)
            if (mask[p] >> b & 1) { if (k == 0) return 5 * S * p + b; k--; }
    return -1;
}
static int random_action(const uint32_t mask[6], uint32_t word)
{
    uint64_t m[6];
    for (int p = 0; p < 6; p++) m[p] = mask[p];
    return random_action64(m, word);
}

/* ------------------------------------------------------------------ exported entry points - */

void ao_init(int32_t *rec, int players, int pool, int first_player)
{
    ao_game g; memset(&g, 0, sizeof(g));
    init_game(&g, players, pool, first_player);
    to_record(&g, rec, players);
}

void ao_new_round(int32_t *rec, int players, int pool, const int8_t *draws)
{
    ao_game g; from_record(&g, rec, players);
    new_round_injected(&g, pool, draws);
    to_record(&g, rec, players);
}

void ao_reset_philox(int32_t *rec, int players, int pool, int first_rule, uint64_t seed, uint32_t gid)
{
    ao_game g; from_record(&g, rec, players);
    reset_philox(&g, players, pool, first_rule, seed, gid);
    to_record(&g, rec, players);
}

void ao_move(int32_t *rec, int players, int pool, int d, int c, int p)
{
    ao_game g; from_record(&g, rec, players);
    move(&g, pool, d, c, p);
    to_record(&g, rec, players);
}

int ao_is_legal_move(const int32_t *rec, int players, int d, int c, int p)
{
    ao_game g; from_record(&g, rec, players);
    return is_legal_move(&g, d, c, p);
}

void ao_next_player(int32_t *rec, int players)
{
    ao_game g; from_record(&g, rec, players);
    next_player(&g);
    to_record(&g, rec, players);
}

int ao_is_end_of_round(const int32_t *rec, int players)
{
    ao_game g; from_record(&g, rec, players);
    return is_end_of_round(&g);
}

int ao_is_end_of_game(const int32_t *rec, int players)
{
    ao_game g; from_record(&g, rec, players);
    return is_end_of_game(&g);
}

void ao_count_score(int32_t *rec, int players, int pool)
{
    ao_game g; from_record(&g, rec, players);
    count_score(&g, pool);
    to_record(&g, rec, players);
}

/* game_runner.py:48-50: score after a count_score on a copy; the record is not modified */
void ao_score_preview(const int32_t *rec, int players, int pool, int32_t *score_out)
{
    ao_game g; from_record(&g, rec, players);
    count_score(&g, pool);
    for (int p = 0; p < players; p++) score_out[p] = g.score[p];
}

void ao_legal_mask(const int32_t *rec, int players, uint32_t *mask6)
{
    ao_game g; from_record(&g, rec, players);
    legal_mask(&g, mask6);
}

/* the same with 64-bit words: needed for 7 / 9 displays (40 / 50 bits per word) */
void ao_legal_mask64(const int32_t *rec, int players, uint64_t *mask6)
{
    ao_game g; from_record(&g, rec, players);
    legal_mask64(&g, mask6);
}

int ao_step(int32_t *rec, int players, int pool, int action, const int8_t *draws, uint64_t seed, uint32_t gid)
{
    ao_game g; from_record(&g, rec, players);
    int rc;
    const int S = g_F + 1;
    if (action < 0 || action >= 30 * S) rc = g.end_of_game ? -2 : -1;
    else rc = step(&g, pool, action % S, (action / S) % 5, action / (5 * S), draws, seed, gid);
    if (rc == 0) to_record(&g, rec, players);
    return rc;
}

/* SPEC of azb_opponent_random = GameRunner.step after the agent's move (game_runner.py:46-52) with a
 * random-agent opponent on the Philox ACTION words; returns score[0]-score[1] after count_score on a copy. */
int ao_opponent_random(int32_t *rec, int players, int pool, uint64_t seed, uint32_t gid, int require_two, uint32_t *mask6)
{
    ao_game g; from_record(&g, rec, players);
    for (;;) {
        legal_mask(&g, mask6);
        if (g.end_of_game) break;                                        /* game_runner.py:46 "and not is_end_of_game()" */
        int n_valid = 0;
        for (int p = 0; p < 6; p++) n_valid += __builtin_popcount(mask6[p]);
        if (g.current_player == 1 && (!require_two || n_valid >= 2)) break;
        if (n_valid == 0) { g.status |= ST_STUCK; break; }
        uint32_t w[4];
        draw_words(seed, gid, g.total_steps >> 2, PURPOSE_ACTION, 0, w);
        int a = random_action(mask6, w[g.total_steps & 3]);
        step(&g, pool, a % 6, (a / 6) % 5, a / 30, NULL, seed, gid);     /* opponent_move, game_runner.py:37-42 */
    }
    to_record(&g, rec, players);
    ao_game cp = g;                                                      /* game_runner.py:48 deepcopy */
    count_score(&cp, pool);                                              /* :49 */
    return cp.score[0] - cp.score[1];                                    /* :50 */
}

int ao_random_action(const uint32_t *mask6, uint32_t word) { return random_action(mask6, word); }

/* GameRunner.get_state (game_runner.py:56-72): displays (25) | centre (6) | pattern_lines[order] (25P) |
 * walls[order] (25P) | floors[order] (P) | score[order] (P) | next first player seen from `perspective` (1);
 * order = [perspective] + the other seats ascending (:57).  obs has 32 + 52P entries. */
void ao_observe(const int32_t *rec, int players, int perspective, int32_t *obs)
{
    ao_game g; from_record(&g, rec, players);
    int P = players, order[MAXP], n = 0, k = 0;
    if (perspective < 0) perspective = seat(&g);                       /* opponent_move, game_runner.py:38 */
    order[n++] = perspective;
    for (int p = 0; p < P; p++) if (p != perspective) order[n++] = p;
    for (int i = 0; i < 5; i++) for (int c = 0; c < 5; c++) obs[k++] = g.displays[i][c];
    for (int c = 0; c < 6; c++) obs[k++] = g.center[c];
    for (int s = 0; s < P; s++) for (int r = 0; r < 5; r++) for (int c = 0; c < 5; c++) obs[k++] = g.pattern_lines[order[s]][r][c];
    for (int s = 0; s < P; s++) for (int r = 0; r < 5; r++) for (int c = 0; c < 5; c++) obs[k++] = g.walls[order[s]][r][c] != 0;
    for (int s = 0; s < P; s++) obs[k++] = g.floors[order[s]];
    for (int s = 0; s < P; s++) obs[k++] = g.score[order[s]];
    /* :58-61; Python's % is non-negative */
    obs[k++] = g.next_first_player > 0 ? (((g.next_first_player - 1 - perspective) % P + P) % P) + 1 : 0;
}

/* The loop test of GameRunner.step / GameRunner.reset as a function of the state alone:
 *   step  (game_runner.py:46): (current_player != 1 or #valid < 2) and not is_end_of_game()   [require_two != 0]
 *   reset (game_runner.py:84): current_player != 1                                            [require_two == 0]
 * (the reference's is_end_of_game() recomputes the test from the walls, azul.py:184-191).  1 = the opponent moves next. */
int ao_runner_continues(const int32_t *rec, int players, int require_two)
{
    ao_game g; from_record(&g, rec, players);
    if (!require_two) return g.current_player != 1;
    uint32_t m[6]; legal_mask(&g, m);
    int n_valid = 0;
    for (int p = 0; p < 6; p++) n_valid += __builtin_popcount(m[p]);
    return (g.current_player != 1 || n_valid < 2) && !is_end_of_game(&g);
}

/* Azul.get_statistics (azul.py:314-315) as ten integers: player_score, opponent_score, rounds,
 * first_player_stats[0], sum(first_player_stats), -floor_penalty[0], max_combo[0], completed rows,
 * completed columns, completed colours of seat 0 (percent_first_player = 100 * [3] / [4], win = [0] > [1]) */
void ao_statistics(const int32_t *rec, int players, int32_t *out)
{
    ao_game g; from_record(&g, rec, players);
    int fps = 0;
    for (int p = 0; p < players; p++) fps += g.first_player_stats[p];
    out[0] = g.score[0]; out[1] = g.score[1]; out[2] = g.turn_counter; out[3] = g.first_player_stats[0]; out[4] = fps;
    out[5] = -g.floor_penalty[0]; out[6] = g.max_combo[0]; out[7] = g.completed_lines[0][0];
    out[8] = g.completed_lines[0][2]; out[9] = g.completed_lines[0][1];
}

/* SPEC of azb_rollout_random: K env steps per slot, random agents on every seat, auto-reset.
 *   per step:  mask; if empty -> STUCK: count, reset, recompute mask
 *              action word = Philox(gid, T>>2, ACTION, 0)[T&3]
 *              step (Philox refill); on game end: add the finished game to the counters, reset.
 * counters[16] (int64, accumulated):
 *   0 steps, 1 games finished, 2 rounds started (new_round calls incl. resets), 3 sum score seat 0,
 *   4 sum score seat 1, 5 games with score0 > score1, 6 stuck aborts, 7 games that ran the bag empty,
 *   8 sum turn_counter of finished games, 9 sum -floor_penalty seat 0, 10 sum max_combo seat 0,
 *   11 completed rows seat 0, 12 completed columns seat 0, 13 completed colours seat 0,
 *   14 sum first_player_stats seat 0, 15 sum of all seats' scores.                             */
static void rollout_one(ao_game *g, int P, int pool, int first_rule, uint64_t seed, uint32_t gid,
                        int k_steps, int64_t *cnt)
{
    const int S = g_F + 1;
    for (int i = 0; i < k_steps; i++) {
        uint64_t mask[6];
        if (g->end_of_game) { reset_philox(g, P, pool, first_rule, seed, gid); cnt[2]++; }
        legal_mask64(g, mask);
        if ((mask[0] | mask[1] | mask[2] | mask[3] | mask[4] | mask[5]) == 0) {
            cnt[6]++;
            reset_philox(g, P, pool, first_rule, seed, gid); cnt[2]++;
            legal_mask64(g, mask);
        }
        uint32_t w[4];
        draw_words(seed, gid, g->total_steps >> 2, PURPOSE_ACTION, 0, w);
        int a = random_action64(mask, w[g->total_steps & 3]);
        int turn_before = g->turn_counter, bag_before = g->status & ST_BAG_EMPTY;
        step(g, pool, a % S, (a / S) % 5, a / (5 * S), NULL, seed, gid);
        cnt[0]++;
        if (g->turn_counter != turn_before) cnt[2]++;
        if (!bag_before && (g->status & ST_BAG_EMPTY)) cnt[7]++;
        if (g->end_of_game) {
            cnt[1]++;
            cnt[3] += g->score[0]; cnt[4] += g->score[1];
            cnt[5] += g->score[0] > g->score[1];
            cnt[8] += g->turn_counter;
            cnt[9] += -g->floor_penalty[0];
            cnt[10] += g->max_combo[0];
            cnt[11] += g->completed_lines[0][0];
            cnt[12] += g->completed_lines[0][2];
            cnt[13] += g->completed_lines[0][1];
            cnt[14] += g->first_player_stats[0];
            for (int p = 0; p < P; p++) cnt[15] += g->score[p];
            reset_philox(g, P, pool, first_rule, seed, gid); cnt[2]++;
        }
    }
}

void ao_rollout_random(int32_t *recs, int64_t n, int players, int pool, int first_rule,
                       uint64_t seed, uint32_t gid0, int k_steps, int64_t *counters)
{
    int U = ao_record_size(players);
    for (int64_t i = 0; i < n; i++) {
        ao_game g; from_record(&g, recs + i * U, players);
        rollout_one(&g, players, pool, first_rule, seed, gid0 + (uint32_t)i, k_steps, counters);
        to_record(&g, recs + i * U, players);
    }
}

typedef struct {
    int32_t *recs; int64_t lo, hi; int players, pool, first_rule; uint64_t seed; uint32_t gid0;
    int k_steps; int64_t cnt[16];
} ao_job;

static void *rollout_thread(void *arg)
{
    ao_job *j = (ao_job *)arg;
    int U = ao_record_size(j->players);
    for (int64_t i = j->lo; i < j->hi; i++) {
        ao_game g; from_record(&g, j->recs + i * U, j->players);
        rollout_one(&g, j->players, j->pool, j->first_rule, j->seed, j->gid0 + (uint32_t)i, j->k_steps, j->cnt);
        to_record(&g, j->recs + i * U, j->players);
    }
    return NULL;
}

/* the same rollout on n_threads host threads (games are independent): CPU baseline leg */
void ao_rollout_random_mt(int32_t *recs, int64_t n, int players, int pool, int first_rule,
                          uint64_t seed, uint32_t gid0, int k_steps, int64_t *counters, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; ao_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        ao_job *j = &jobs[t];
        memset(j, 0, sizeof(*j));
        j->recs = recs; j->lo = n * t / n_threads; j->hi = n * (t + 1) / n_threads;
        j->players = players; j->pool = pool; j->first_rule = first_rule; j->seed = seed;
        j->gid0 = gid0; j->k_steps = k_steps;
        pthread_create(&th[t], NULL, rollout_thread, j);
    }
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        for (int k = 0; k < 16; k++) counters[k] += jobs[t].cnt[k];
    }
}
