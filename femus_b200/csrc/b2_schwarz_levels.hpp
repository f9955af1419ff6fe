// Host-only helper of the element-block smoother (b2_schwarz.cu; also used by the CPU emulator harness of the tests):
// dependency levels of the rows of every block in its lower (dir 0: forward sweeps, ILU elimination) and upper (dir 1:
// backward sweeps) triangular in-block pattern.  Row i of a block depends on the block's rows j whose dof is a column
// of row i below (above) the diagonal; level(i) = 1 + max level(j).  The backward SSOR sweep also READS the columns below
// the diagonal, which must still hold their forward values: in dir 1 a row j < i that row i reads is therefore placed
// after row i as well (on a structurally symmetric pattern that is already implied; a non-symmetric coupling table,
// which BuildSystemSparsity accepts, needs it).  Output per direction: ptr[nblocks+1] into off,
// off = per block its nlev+1 offsets into the block's row list, rows[blk_ptr[b] + t] = local rows sorted by level
// (inside a level in the sweep's own order).  Returns the longest chain.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

inline int64_t b2_schwarz_row_level_schedule(int64_t nblocks, const int64_t* bp, const int32_t* bd, const int64_t* rp, const int32_t* col,
                                             std::vector<int64_t> ptr[2], std::vector<int32_t> off[2], std::vector<int32_t> rows[2]) {
  int64_t max_levels = 0;
  std::vector<int32_t> lvl, cnt, after;
  for (int dir = 0; dir < 2; dir++) {
    ptr[dir].assign(1, 0);
    off[dir].clear();
    rows[dir].assign((size_t)bp[nblocks], 0);
  }
  for (int64_t b = 0; b < nblocks; b++) {
    const int32_t* D = bd + bp[b];
    const int m = (int)(bp[b + 1] - bp[b]);
    for (int dir = 0; dir < 2; dir++) {
      lvl.assign(m, 0);
      after.assign(m, 0);      // dir 1: lowest level a row may take because a later-numbered row reads its forward value
      int nlev = 0;
      for (int ii = 0; ii < m; ii++) {
        const int i = dir ? m - 1 - ii : ii;
        const int64_t r = D[i];
        int lv = after[i];
        for (int64_t q = rp[r]; q < rp[r + 1]; q++) {
          const int32_t cc = col[q];
          if (dir ? cc <= r : cc >= r) continue;
          const int32_t* it = std::lower_bound(D, D + m, cc);
          if (it != D + m && *it == cc) lv = std::max(lv, lvl[it - D] + 1);
        }
        lvl[i] = lv;
        nlev = std::max(nlev, lv + 1);
        if (dir)
          for (int64_t q = rp[r]; q < rp[r + 1] && col[q] < r; q++) {
            const int32_t* it = std::lower_bound(D, D + m, col[q]);
            if (it != D + m && *it == col[q]) after[it - D] = std::max(after[it - D], lv + 1);
          }
      }
      cnt.assign(nlev + 1, 0);
      for (int i = 0; i < m; i++) cnt[lvl[i] + 1]++;
      for (int l = 0; l < nlev; l++) cnt[l + 1] += cnt[l];
      off[dir].insert(off[dir].end(), cnt.begin(), cnt.end());
      ptr[dir].push_back((int64_t)off[dir].size());
      std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
      for (int ii = 0; ii < m; ii++) {
        const int i = dir ? m - 1 - ii : ii;
        rows[dir][bp[b] + fill[lvl[i]]++] = i;
      }
      max_levels = std::max<int64_t>(max_levels, nlev);
    }
  }
  return max_levels;
}
