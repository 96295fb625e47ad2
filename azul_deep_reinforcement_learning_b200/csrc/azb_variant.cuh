// azb_variant.cuh -- the opt-in rule variant "factory count by player count" (SURVEY §8f rank 4): 5 / 7 / 9 factory
// displays for 2 / 3 / 4 players as in the board game, i.e. 180 / 240 / 300 actions.  The reference never implemented
// it (azul.py:72 is a TODO; azul.py:19 always builds five displays), so there is no reference behaviour to be exact
// against: the spec is the reference's rules with `range(5)` over the displays replaced by `range(F)`, restated in the test
// suite's C checker (run with 7 / 9 displays) and checked against it.  With F = 5 this code must -- and is tested to --
// reproduce the default, reference-pinned engine bit for bit.  The default path (azb_rules.cuh's Game<P>) is untouched.
//
// Packed state, Wv(P,F) = ceil(F/2) + 5 + 5P words per game, structure-of-arrays like the default layout:
//   0 .. DW-1  DISP   display i (0-based) in word i/2 at bit 15*(i&1) + 3c: count of colour c (0..4)
//   DW         CEN    centre: colour c at bits [5c+4:5c] (up to 27 tiles of a colour with nine displays), token at bit 25
//   DW+1       MISC   [8:6] current_player [11:9] next_first_player [12] end_of_game [15:13] sticky status [27:16] turn_counter
//   DW+2..4    BOX, LID, STEPS as in the default layout
//   per player the same five words as the default layout (PAT, WALL, SCF, STA, STB): scoring is shared code.
// Action a = d + S*c + 5*S*p with S = F + 1 sources (game_runner.py:102-103 with 6 -> S); legal mask = six 64-bit words,
// word p bit (d + S*c).
#pragma once
#include "azb_rules.cuh"

namespace azb {

template <int P, int F>
struct GameV {
    static constexpr int PLAYERS = P, FACT = F, S = F + 1, DW = (F + 1) / 2, WORDS = DW + 5 + 5 * P, N_ACT = 30 * S;
    uint32_t disp[DW], cen, misc, box, lid, steps;
    uint32_t pat[P], wall[P], scf[P], sta[P], stb[P];

    AZB_M void load(const uint32_t* __restrict__ s, int64_t stride, int64_t g)
    {
#pragma unroll
        for (int w = 0; w < DW; w++) disp[w] = s[w * stride + g];
        cen = s[DW * stride + g]; misc = s[(DW + 1) * stride + g]; box = s[(DW + 2) * stride + g];
        lid = s[(DW + 3) * stride + g]; steps = s[(DW + 4) * stride + g];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int b = DW + 5 + 5 * p;
            pat[p] = s[b * stride + g]; wall[p] = s[(b + 1) * stride + g]; scf[p] = s[(b + 2) * stride + g];
            sta[p] = s[(b + 3) * stride + g]; stb[p] = s[(b + 4) * stride + g];
        }
    }
    AZB_M void store(uint32_t* __restrict__ s, int64_t stride, int64_t g) const
    {
#pragma unroll
        for (int w = 0; w < DW; w++) s[w * stride + g] = disp[w];
        s[DW * stride + g] = cen; s[(DW + 1) * stride + g] = misc; s[(DW + 2) * stride + g] = box;
        s[(DW + 3) * stride + g] = lid; s[(DW + 4) * stride + g] = steps;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int b = DW + 5 + 5 * p;
            s[b * stride + g] = pat[p]; s[(b + 1) * stride + g] = wall[p]; s[(b + 2) * stride + g] = scf[p];
            s[(b + 3) * stride + g] = sta[p]; s[(b + 4) * stride + g] = stb[p];
        }
    }

    AZB_M uint32_t current_player() const { return (misc >> 6) & 7u; }
    AZB_M uint32_t next_first_player() const { return (misc >> 9) & 7u; }
    AZB_M bool ended() const { return (misc >> 12) & 1u; }
    AZB_M uint32_t status() const { return ((misc >> 13) & 7u) << 2; }
    AZB_M uint32_t turn_counter() const { return (misc >> 16) & 0xFFFu; }
    AZB_M void set_current_player(uint32_t v) { misc = (misc & ~(7u << 6)) | (v << 6); }
    AZB_M void set_next_first_player(uint32_t v) { misc = (misc & ~(7u << 9)) | (v << 9); }
    AZB_M void add_status(uint32_t bits) { misc |= ((bits >> 2) & 7u) << 13; }
    AZB_M int seat() const { uint32_t c = current_player(); return c ? (int)c - 1 : P - 1; }
    AZB_M uint32_t sel(const uint32_t (&a)[P], int s) const
    {
        uint32_t v = a[0];
#pragma unroll
        for (int p = 1; p < P; p++) v = (s == p) ? a[p] : v;
        return v;
    }
    AZB_M void put(uint32_t (&a)[P], int s, uint32_t v)
    {
#pragma unroll
        for (int p = 0; p < P; p++) a[p] = (s == p) ? v : a[p];
    }
    // 15 bits of display i (five 3-bit colour counts); i is a run-time index: predicated selects, no local memory
    AZB_M uint32_t display_bits(int i) const
    {
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < DW; w++) v = (w == (i >> 1)) ? disp[w] : v;
        return (v >> (15 * (i & 1))) & 0x7FFFu;
    }
    AZB_M void display_add(int i, uint32_t bits15)
    {
#pragma unroll
        for (int w = 0; w < DW; w++) disp[w] += (w == (i >> 1)) ? bits15 << (15 * (i & 1)) : 0u;
    }
    AZB_M void display_clear(int i)
    {
#pragma unroll
        for (int w = 0; w < DW; w++) disp[w] &= (w == (i >> 1)) ? ~(0x7FFFu << (15 * (i & 1))) : 0xFFFFFFFFu;
    }
};

// ---- check_all_valid (game_runner.py:113-117) over is_legal_move (azul.py:162-176) ----
template <int P, int F>
AZB_HD void legal_mask_v(const GameV<P, F>& g, uint64_t m[6])
{
    constexpr int S = F + 1;
    uint64_t src = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        src |= (uint64_t)(((g.cen >> (5 * c)) & 31u) ? 1u : 0u) << (S * c);                       // d = 0: the centre
#pragma unroll
        for (int i = 0; i < F; i++)
            src |= (uint64_t)(((g.disp[i >> 1] >> (15 * (i & 1) + 3 * c)) & 7u) ? 1u : 0u) << (i + 1 + S * c);
    }
    const int s = g.seat();
    const uint32_t pat = g.sel(g.pat, s), wall = g.sel(g.wall, s);
    m[0] = src;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint32_t cnt = (pat >> (6 * r + 3)) & 7u, col = (pat >> (6 * r)) & 7u;
        const uint32_t allowed = (cnt ? (1u << col) : 31u) & ~(wall >> (5 * r)) & 31u;          // azul.py:171-175
        uint64_t cm = 0;
#pragma unroll
        for (int c = 0; c < 5; c++) cm |= ((allowed >> c) & 1u) ? (((1ull << S) - 1ull) << (S * c)) : 0ull;
        m[r + 1] = src & cm;
    }
}

// ---- move (azul.py:118-161); no legality check, like the reference ----
template <int P, int F, int POOL>
AZB_HD void apply_move_v(GameV<P, F>& g, uint32_t action)
{
    constexpr uint32_t S = F + 1;
    const uint32_t p = action / (5u * S), b = action - 5u * S * p, c = b / S, d = b - S * c;
    const int s = g.seat();
    uint32_t n, tok = 0u;
    if (d != 0u) {                                                    // azul.py:125-133
        const uint32_t bits = g.display_bits((int)d - 1);
        n = (bits >> (3u * c)) & 7u;
        g.display_clear((int)d - 1);
#pragma unroll
        for (int k = 0; k < 5; k++) g.cen += ((uint32_t)k != c ? (bits >> (3 * k)) & 7u : 0u) << (5 * k);
    } else {                                                          // azul.py:134-143
        n = (g.cen >> (5u * c)) & 31u;
        g.cen &= ~(31u << (5u * c));
        tok = (g.cen >> 25) & 1u;
        g.cen &= ~(1u << 25);
        if (tok) g.set_next_first_player(g.current_player());
    }
    // azul.py:145-161: fill row p-1 up to its capacity p, the rest (everything when p = 0) falls to the floor
    const bool to_row = p != 0u;
    const uint32_t pat = g.sel(g.pat, s);
    const uint32_t sh = to_row ? 6u * (p - 1u) : 0u;
    const uint32_t cnt = (pat >> (sh + 3u)) & 7u;
    const uint32_t room = to_row ? p - cnt : 0u;
    const uint32_t placed = n < room ? n : room;
    const uint32_t to_floor = n - placed;
    const uint32_t newcnt = cnt + placed;
    const uint32_t newpat = (pat & ~(63u << sh)) | ((newcnt ? (c | (newcnt << 3)) : 0u) << sh);
    g.put(g.pat, s, to_row ? newpat : pat);
    g.put(g.scf, s, floor_add(g.sel(g.scf, s), tok + to_floor));
    if (POOL == POOL_LID) g.lid += to_floor << (6u * c);
}

// azul.py:182-183
template <int P, int F>
AZB_HD bool is_end_of_round_v(const GameV<P, F>& g)
{
    uint32_t any = g.cen;
#pragma unroll
    for (int w = 0; w < GameV<P, F>::DW; w++) any |= g.disp[w];
    return any == 0u;
}

// azul.py:64-73
template <int P, int F>
AZB_HD void new_round_header_v(GameV<P, F>& g)
{
    const uint32_t nf = g.next_first_player();
    g.set_current_player(nf);
    const int s = nf ? (int)nf - 1 : P - 1;
#pragma unroll
    for (int p = 0; p < P; p++) g.sta[p] += (s == p) ? 1u : 0u;
    g.misc = (g.misc & ~(0xFFFu << 16)) | (((g.turn_counter() + 1u) & 0xFFFu) << 16);
    g.set_next_first_player(0u);
    g.cen = 1u << 25;
#pragma unroll
    for (int w = 0; w < GameV<P, F>::DW; w++) g.disp[w] = 0u;
}

// azul.py:64-89 with the Philox draw schedule generalised to F displays: Random pool -- display i draws its four colours
// from word i of the call sequence; Lid pool -- display i draws two colours from each of words 2i, 2i + 1
template <int P, int F, int POOL>
AZB_HD void new_round_philox_v(GameV<P, F>& g, const Philox& rng, uint32_t gid, uint32_t purpose)
{
    new_round_header_v(g);
    constexpr uint32_t CALLS = POOL == POOL_LID ? (2u * F + 3u) / 4u : (F + 3u) / 4u;
    BoxRegs B;
    if (POOL == POOL_LID) { if (!B.unpack(g.box)) g.add_status(ST_BAD_IMPORT); }
    AZB_ROLLED
    for (uint32_t j = 0; j < CALLS; j++) {
        uint32_t w[4];
        rng(gid, g.steps, purpose, j, w);
        AZB_ROLLED
        for (uint32_t q = 0; q < 4u; q++) {
            uint32_t x = q == 0u ? w[0] : q == 1u ? w[1] : q == 2u ? w[2] : w[3];
            if (POOL == POOL_RANDOM) {
                const uint32_t i = 4u * j + q;
                if (i >= (uint32_t)F) break;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t c = mulhi(x, 5u);
                    x *= 5u;
                    g.display_add((int)i, 1u << (3u * c));
                }
            } else {
                const uint32_t hidx = 4u * j + q, i = hidx >> 1;
                if (i >= (uint32_t)F) break;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int c = lid_draw(g, B, x);
                    if (c >= 0) g.display_add((int)i, 1u << (3 * c));
                }
            }
        }
    }
    if (POOL == POOL_LID) g.box = B.pack();
}

template <int P, int F, int POOL, typename DrawFn>
AZB_HD void new_round_injected_v(GameV<P, F>& g, DrawFn draw)
{
    new_round_header_v(g);
#pragma unroll 1
    for (int k = 0; k < 4 * F; k++) {
        const int c = draw(k);
        if (c < 0 || c > 4) continue;
        if (POOL == POOL_LID) {
            const uint32_t tot = (g.box & 63u) + ((g.box >> 6) & 63u) + ((g.box >> 12) & 63u) + ((g.box >> 18) & 63u) + ((g.box >> 24) & 63u);
            if (tot == 0u) { g.box = g.lid; g.lid = 0u; }
            if (((g.box >> (6 * c)) & 63u) == 0u) { g.add_status(ST_BAG_EMPTY); continue; }
            g.box -= 1u << (6 * c);
        }
        g.display_add(k / 4, 1u << (3 * c));
    }
}

template <int P, int F, int POOL>
AZB_HD void init_game_v(GameV<P, F>& g, uint32_t first_player)
{
#pragma unroll
    for (int w = 0; w < GameV<P, F>::DW; w++) g.disp[w] = 0u;
    g.cen = 0u;
    g.misc = first_player << 9;
    g.box = (POOL == POOL_LID) ? (20u | 20u << 6 | 20u << 12 | 20u << 18 | 20u << 24) : 0u;
    g.lid = 0u;
#pragma unroll
    for (int p = 0; p < P; p++) { g.pat[p] = g.wall[p] = g.scf[p] = g.sta[p] = g.stb[p] = 0u; }
}

template <int P, int F, int POOL>
AZB_HD void reset_game_v(GameV<P, F>& g, const Philox& rng, uint32_t gid, int first_rule)
{
    uint32_t first = (uint32_t)first_rule;
    if (first_rule == 0) {
        uint32_t w[4];
        rng(gid, g.steps, PURPOSE_FIRST, 0u, w);
        first = 1u + mulhi(w[0], (uint32_t)P);
    }
    init_game_v<P, F, POOL>(g, first);
    new_round_philox_v<P, F, POOL>(g, rng, gid, PURPOSE_RESET_REFILL);
}

// azul.py:296-313 after the legality / ended checks; refill(g) supplies the next round's tiles
template <int P, int F, int POOL, typename RefillFn>
AZB_HD bool advance_v(GameV<P, F>& g, uint32_t action, RefillFn refill)
{
    apply_move_v<P, F, POOL>(g, action);
    g.steps += 1u;
    if (is_end_of_round_v(g)) {
        count_score_g<POOL>(g);
        if (is_end_of_game(g)) { g.misc |= 1u << 12; return true; }
        refill(g);
    } else {
        next_player(g);
    }
    return false;
}

template <int F>
AZB_HD bool action_is_legal_v(const uint64_t m[6], uint32_t action)
{
    constexpr uint32_t S = F + 1;
    if (action >= 30u * S) return false;
    const uint32_t p = action / (5u * S), b = action - 5u * S * p;
    uint64_t w = m[0];
#pragma unroll
    for (int i = 1; i < 6; i++) w = (p == (uint32_t)i) ? m[i] : w;
    return (w >> b) & 1ull;
}

AZB_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// position of the k-th (0-based) set bit of a 64-bit word; k < popc64(m)
AZB_HD uint32_t select_bit64(uint64_t m, uint32_t k)
{
    const uint32_t lo = (uint32_t)m, nlo = (uint32_t)popc((uint32_t)m);
    return k < nlo ? select_bit(lo, k) : 32u + select_bit((uint32_t)(m >> 32), k - nlo);
}

// the integer random agent (game_runner.py:87-97) on 64-bit mask words: floor actions weigh 1, every other legal action
// 100; heavy actions (words 1..5, ascending action index) first.  Returns N_ACT when no action is legal.
template <int F>
AZB_HD uint32_t random_action_v(const uint64_t m[6], uint32_t word)
{
    constexpr uint32_t S = F + 1;
    uint32_t e[6];
    e[0] = 0u;
#pragma unroll
    for (int i = 1; i < 6; i++) e[i] = e[i - 1] + (uint32_t)popc64(m[i]);
    const uint32_t n_hi = e[5], n0 = (uint32_t)popc64(m[0]);
    const uint32_t total = 100u * n_hi + n0;
    if (total == 0u) return 30u * S;
    const uint32_t r = mulhi(word, total);
    if (r < 100u * n_hi) {
        const uint32_t kh = r / 100u;
        uint32_t i = 1u;
#pragma unroll
        for (int t = 1; t < 5; t++) i += kh >= e[t] ? 1u : 0u;
        uint64_t w = m[1];
        uint32_t before = 0u;
#pragma unroll
        for (int t = 2; t < 6; t++) { w = (i == (uint32_t)t) ? m[t] : w; before = (i == (uint32_t)t) ? e[t - 1] : before; }
        return 5u * S * i + select_bit64(w, kh - before);
    }
    return select_bit64(m[0], r - 100u * n_hi);
}

// ---- unpacked record <-> packed game: the default record layout (layout.py) with F displays in front ----
template <int P, int F, typename Rd>
AZB_HD bool import_record_v(GameV<P, F>& g, Rd rd)
{
    bool ok = true;
#pragma unroll
    for (int w = 0; w < GameV<P, F>::DW; w++) g.disp[w] = 0u;
    g.cen = 0u; g.misc = 0u; g.box = g.lid = 0u;
    for (int i = 0; i < F; i++)
        for (int c = 0; c < 5; c++) {
            const int32_t n = rd(i * 5 + c);
            ok &= (n >= 0 && n <= 7);
            g.display_add(i, ((uint32_t)n & 7u) << (3 * c));
        }
    const int o_c = 5 * F;
    for (int c = 0; c < 5; c++) {
        const int32_t n = rd(o_c + c);
        ok &= (n >= 0 && n <= 31);
        g.cen |= ((uint32_t)n & 31u) << (5 * c);
    }
    { const int32_t t = rd(o_c + 5); ok &= (t == 0 || t == 1); g.cen |= (uint32_t)(t & 1) << 25; }
    const int o_pat = o_c + 6, o_wall = o_pat + 25 * P, o_fl = o_pat + 50 * P, o_sc = o_pat + 51 * P, o_s = o_pat + 52 * P;
    for (int p = 0; p < P; p++) {
        uint32_t pat = 0, wall = 0;
        for (int r = 0; r < 5; r++) {
            int colours = 0;
            for (int c = 0; c < 5; c++) {
                const int32_t n = rd(o_pat + 25 * p + 5 * r + c);
                ok &= (n >= 0 && n <= 7);
                if (n != 0) { colours++; pat |= ((uint32_t)c | ((uint32_t)n << 3)) << (6 * r); }
                if (rd(o_wall + 25 * p + 5 * r + c) != 0) wall |= 1u << (5 * r + c);
            }
            if (colours > 1) { ok = false; pat &= ~(63u << (6 * r)); }
        }
        const int32_t fl = rd(o_fl + p), sc = rd(o_sc + p);
        ok &= (fl >= 0 && fl <= 7 && sc >= 0 && sc <= 0xFFFF);
        g.pat[p] = pat; g.wall[p] = wall;
        g.scf[p] = ((uint32_t)sc & 0xFFFFu) | (((uint32_t)fl & 7u) << 16);
        const int32_t fps = rd(o_s + 15 + p), fpen = -rd(o_s + 15 + P + p), mc = rd(o_s + 15 + 2 * P + p);
        ok &= (fps >= 0 && fps <= 0xFFF && fpen >= 0 && fpen <= 0xFFFF && mc >= 0 && mc <= 15);
        g.sta[p] = ((uint32_t)fps & 0xFFFu) | (((uint32_t)fpen & 0xFFFFu) << 12) | (((uint32_t)mc & 15u) << 28);
        const int32_t cr = rd(o_s + 15 + 3 * P + 3 * p), cc = rd(o_s + 15 + 3 * P + 3 * p + 1), ck = rd(o_s + 15 + 3 * P + 3 * p + 2);
        ok &= (cr >= 0 && cr <= 255 && cc >= 0 && cc <= 255 && ck >= 0 && ck <= 255);
        g.stb[p] = ((uint32_t)cr & 255u) | (((uint32_t)cc & 255u) << 8) | (((uint32_t)ck & 255u) << 16);
    }
    const int32_t cur = rd(o_s + 0), nf = rd(o_s + 1), eog = rd(o_s + 3), turn = rd(o_s + 4);
    ok &= (cur >= 0 && cur <= P && nf >= 0 && nf <= P && turn >= 0 && turn <= 0xFFF && rd(o_s + 2) == P);
    g.misc |= ((uint32_t)cur & 7u) << 6 | ((uint32_t)nf & 7u) << 9 | (eog ? 1u << 12 : 0u) | ((uint32_t)turn & 0xFFFu) << 16;
    for (int c = 0; c < 5; c++) {
        const int32_t b = rd(o_s + 5 + c), l = rd(o_s + 10 + c);
        ok &= (b >= 0 && b <= 63 && l >= 0 && l <= 63);
        g.box |= ((uint32_t)b & 63u) << (6 * c); g.lid |= ((uint32_t)l & 63u) << (6 * c);
    }
    g.steps = (uint32_t)rd(o_s + 15 + 6 * P);
    g.add_status((uint32_t)rd(o_s + 16 + 6 * P) & (ST_STUCK | ST_BAG_EMPTY | ST_BAD_IMPORT));
    if (!ok) g.add_status(ST_BAD_IMPORT);
    return ok;
}

template <int P, int F, typename Wr>
AZB_HD void export_record_v(const GameV<P, F>& g, Wr wr)
{
    for (int i = 0; i < F; i++)
        for (int c = 0; c < 5; c++) wr(i * 5 + c, (int32_t)((g.display_bits(i) >> (3 * c)) & 7u));
    const int o_c = 5 * F;
    for (int c = 0; c < 5; c++) wr(o_c + c, (int32_t)((g.cen >> (5 * c)) & 31u));
    wr(o_c + 5, (int32_t)((g.cen >> 25) & 1u));
    const int o_pat = o_c + 6, o_wall = o_pat + 25 * P, o_fl = o_pat + 50 * P, o_sc = o_pat + 51 * P, o_s = o_pat + 52 * P;
    for (int p = 0; p < P; p++) {
        for (int r = 0; r < 5; r++) {
            const uint32_t cnt = (g.pat[p] >> (6 * r + 3)) & 7u, col = (g.pat[p] >> (6 * r)) & 7u;
            for (int c = 0; c < 5; c++) {
                wr(o_pat + 25 * p + 5 * r + c, (int32_t)((cnt && col == (uint32_t)c) ? cnt : 0u));
                wr(o_wall + 25 * p + 5 * r + c, (int32_t)((g.wall[p] >> (5 * r + c)) & 1u));
            }
        }
        wr(o_fl + p, (int32_t)((g.scf[p] >> 16) & 7u));
        wr(o_sc + p, (int32_t)(g.scf[p] & 0xFFFFu));
        wr(o_s + 15 + p, (int32_t)(g.sta[p] & 0xFFFu));
        wr(o_s + 15 + P + p, -(int32_t)((g.sta[p] >> 12) & 0xFFFFu));
        wr(o_s + 15 + 2 * P + p, (int32_t)(g.sta[p] >> 28));
        wr(o_s + 15 + 3 * P + 3 * p + 0, (int32_t)(g.stb[p] & 255u));
        wr(o_s + 15 + 3 * P + 3 * p + 1, (int32_t)((g.stb[p] >> 8) & 255u));
        wr(o_s + 15 + 3 * P + 3 * p + 2, (int32_t)((g.stb[p] >> 16) & 255u));
    }
    wr(o_s + 0, (int32_t)g.current_player());
    wr(o_s + 1, (int32_t)g.next_first_player());
    wr(o_s + 2, P);
    wr(o_s + 3, (int32_t)g.ended());
    wr(o_s + 4, (int32_t)g.turn_counter());
    for (int c = 0; c < 5; c++) {
        wr(o_s + 5 + c, (int32_t)((g.box >> (6 * c)) & 63u));
        wr(o_s + 10 + c, (int32_t)((g.lid >> (6 * c)) & 63u));
    }
    wr(o_s + 15 + 6 * P, (int32_t)g.steps);
    wr(o_s + 16 + 6 * P, (int32_t)g.status());
}

// ---- K env steps of the random agent with auto-reset: the checker's rollout spec, one game per thread ----
template <int P, int F, int POOL, typename Sink>
AZB_HD void rollout_steps_v(GameV<P, F>& g, const Philox& rng, uint32_t gid, int first_rule, int k_steps, Sink& sink)
{
    uint32_t rounds = 0;
    AZB_ROLLED
    for (int i = 0; i < k_steps; i++) {
        uint64_t m[6];
        if (g.ended()) { reset_game_v<P, F, POOL>(g, rng, gid, first_rule); rounds++; }
        legal_mask_v(g, m);
        if (m[0] == 0ull) {                                           // stuck round: abort the game
            sink.add(6, 1);
            reset_game_v<P, F, POOL>(g, rng, gid, first_rule); rounds++;
            legal_mask_v(g, m);
        }
        uint32_t w[4];
        rng(gid, g.steps >> 2, PURPOSE_ACTION, 0u, w);
        const uint32_t idx = g.steps & 3u;
        const uint32_t word = idx == 0u ? w[0] : idx == 1u ? w[1] : idx == 2u ? w[2] : w[3];
        const uint32_t turn_before = g.turn_counter(), bag_before = g.status() & ST_BAG_EMPTY;
        advance_v<P, F, POOL>(g, random_action_v<F>(m, word),
                              [&](GameV<P, F>& gg) { new_round_philox_v<P, F, POOL>(gg, rng, gid, PURPOSE_REFILL); });
        if (g.turn_counter() != turn_before) rounds++;
        if (!bag_before && (g.status() & ST_BAG_EMPTY)) sink.add(7, 1);
        if (g.ended()) {
            tally_finished(g, sink);
            reset_game_v<P, F, POOL>(g, rng, gid, first_rule); rounds++;
        }
    }
    sink.add(0, (uint32_t)k_steps);
    sink.add(2, rounds);
}

}  // namespace azb
