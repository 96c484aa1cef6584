// Device-memory layout of one ensemble (B trajectories of an n-dimensional SDE) and the POD handed to kernels.
//
// Everything a trajectory owns is stored structure-of-arrays with the trajectory index fastest, so that a warp whose
// 32 lanes hold 32 consecutive trajectories reads and writes 256 contiguous bytes per double field:
//
//   y        [n][B]      state                                   (reference: sde_integrator.state, template:29)
//   t, dt    [B]         time, controller step size              (template:30; _jitcsde.py self.dt)
//   new_y, err [n][B], new_t [B]   last proposal                 (template:31-33) — only the single-step surface uses them
//   q        [DP][B][IT] noise memory (template:15-21) as a bounded ring instead of a list: DP = D rounded up to a power
//                        of two; item d of trajectory b lives in slot (q_head[b] + d) mod DP and is {h, DW[n], DZ[n]}
//                        padded to IT = 2*ceil((1+2n)/2) doubles; item 0 starts at the current time.  Dropping the
//                        consumed items (accept_step) advances the head; a Brownian-bridge insertion moves whichever
//                        side of the insertion point is shorter (usually nothing)
//   q_count, q_cursor, q_head [B]
//   mt       [B][156] uint4   generator state, one contiguous 2496-byte record per trajectory (random_numbers.c:31-38);
//                             positions of different trajectories drift apart (rejections), so records are lane-private
//   mt_chunk [B]         next 4-word chunk to generate (the reference's `pos` / 4, running ahead by the queued pairs)
//   rng_spill [CAP][3][B], rng_level [B]   polar pairs generated but not yet consumed (warp_rng.cuh), kept between launches
//   par      [P] or [P][B]    control parameters (template:34-36)
//   status   [B]; accepted, rejected [B]
//   jump scheduling: next_jump [B], jump_set [B], tape_pos [B], tape taus/xis [B][L]
#pragma once

#include <stdint.h>

namespace jsde {

__host__ __device__ constexpr int noise_item_doubles(int n) { return (1 + 2 * n + 1) & ~1; }

enum { TRAJ_OK = 0, TRAJ_MIN_STEP = 1, TRAJ_NOISE_OVERFLOW = 2 };

struct Controller {
	double atol, rtol, min_step, max_step, dec_thr, inc_thr, safety, max_factor, min_factor;
	double shrink_exp;  // -1/q        (_jitcsde.py:587)
	double grow_exp;    // -1/(q+1)    (_jitcsde.py:592)
};

struct CallTotals {
	unsigned long long accepted, rejected, fresh_draws, jumps;
	long long n_min_step, n_overflow;
	int max_depth;
	// k_integrate with lane recycling: trajectories beyond the grid's first wave wait in EnsembleView::pool[0 .. waiting)
	int waiting;                // shrinks when a lane whose trajectory is done pops the last one
	unsigned long long cursor;  // rotating position of the yield exchanges
	unsigned long long tape_exhausted;  // trajectories that wanted another jump from a tape that had none left
};

struct EnsembleView {
	int64_t B;
	int D;      // noise capacity (items)
	int Dmask;  // physical ring size - 1 (power of two >= D)
	double *y, *t, *dt;
	double *new_y, *err, *new_t;
	double *q;  // [Dmask + 1][B][IT]
	int *q_count, *q_cursor, *q_head;
	uint4 *mt;
	int *mt_chunk;
	double *rng_spill;  // [CAP][3][B] accepted-but-unconsumed pairs (x1, x2, r2) between kernel launches
	int *rng_level;     // [B]
	const double *par;
	int64_t par_stride;  // 0: shared, B: per trajectory
	int *status;
	unsigned long long *accepted, *rejected;
	double *next_jump;
	unsigned char *jump_set;
	long long *tape_pos;
	const double *taus, *xis;
	long long tape_len;
	// counter-based jump variates instead of a tape (philox_jumps.h): generated where they are needed
	int generated_jumps;
	unsigned long long jump_seed;
	long long jump_first;  // global index of trajectory 0 (sharded ensembles draw the same variates as one big one)
	CallTotals *totals;
	int *pool;         // [B] trajectory numbers waiting for a lane (k_integrate, RECYCLE)
	int slice_rounds;  // rounds after which a warp's lanes exchange their trajectories with waiting ones (0: never)
};

}  // namespace jsde
