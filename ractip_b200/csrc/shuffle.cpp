// shuffle.cpp -- host generator of the --zscore shuffle batch.
//
// Replaces the generator half of the loop at reference
// src/ractip.cpp:1636-1643: srandom(seed_); then, per iteration, a
// k-let-preserving shuffle (k=2) of fa1 and of fa2, each drawn from the
// ORIGINAL sequence.  The shuffle follows uShuffle's published algorithm as
// used by the reference (src/ushuffle.c:139-275): a multigraph on the distinct
// (k-1)-lets, a uniformly random arborescence towards the last let by Wilson's
// loop-erased walks, a Fisher-Yates permutation of the remaining out-edges,
// and an Euler walk from the first let.  Random numbers are consumed in the
// same order and with the same `% m` reductions, from glibc's TYPE_3
// additive-feedback generator, so a given --seed yields the reference's exact
// strings.  random_r() with a private 128-byte state reproduces
// srandom()/random() without disturbing the host program's global generator.
//
// The whole batch can be generated up front because nothing in solve()
// consumes random().
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ractip_prob.h"

namespace {

class GlibcRandom {
 public:
  explicit GlibcRandom(unsigned seed) {
    std::memset(&data_, 0, sizeof data_);
    std::memset(state_, 0, sizeof state_);
    initstate_r(seed, state_, sizeof state_, &data_);
  }
  long next() {
    int32_t r;
    random_r(&data_, &r);
    return r;
  }

 private:
  char state_[128];
  struct random_data data_;
};

struct Vertex {
  std::vector<int> out;  // successor vertices, one per occurrence, in sequence order
  int next = 0;          // chosen arborescence edge (slot in `out`)
  bool intree = false;
  int first_pos = 0;     // position of the first occurrence of this let
};

template <class T>
void fisher_yates(T* t, int l, GlibcRandom& rng) {
  for (int i = l - 1; i > 0; i--) {
    int j = (int)(rng.next() % (i + 1));
    T tmp = t[i];
    t[i] = t[j];
    t[j] = tmp;
  }
}

void klet_shuffle(const char* s, char* t, int l, int k, GlibcRandom& rng) {
  if (k >= l) {  // exact copy
    std::memcpy(t, s, l);
    return;
  }
  if (k <= 1) {  // plain permutation
    std::memcpy(t, s, l);
    fisher_yates(t, l, rng);
    return;
  }
  const int n_lets = l - k + 2;  // number of (k-1)-lets
  std::map<std::string, int> ids;
  std::vector<int> let_vertex(n_lets);
  std::vector<Vertex> V;
  for (int i = 0; i < n_lets; i++) {
    std::string let(s + i, k - 1);
    auto it = ids.find(let);
    if (it == ids.end()) {
      it = ids.emplace(let, (int)V.size()).first;
      V.emplace_back();
      V.back().first_pos = i;
    }
    let_vertex[i] = it->second;
  }
  const int root = let_vertex[n_lets - 1];
  for (int i = 0; i + 1 < n_lets; i++) V[let_vertex[i]].out.push_back(let_vertex[i + 1]);

  // Wilson: random arborescence rooted at the last let
  V[root].intree = true;
  for (size_t i = 0; i < V.size(); i++) {
    int u = (int)i;
    while (!V[u].intree) {
      V[u].next = (int)(rng.next() % (long)V[u].out.size());
      u = V[u].out[V[u].next];
    }
    u = (int)i;
    while (!V[u].intree) {
      V[u].intree = true;
      u = V[u].out[V[u].next];
    }
  }
  // the tree edge goes last, the others are permuted
  for (size_t i = 0; i < V.size(); i++) {
    Vertex& u = V[i];
    int n = (int)u.out.size();
    if ((int)i != root) {
      int j = u.out[n - 1];
      u.out[n - 1] = u.out[u.next];
      u.out[u.next] = j;
      fisher_yates(u.out.data(), n - 1, rng);
    } else {
      fisher_yates(u.out.data(), n, rng);
    }
  }
  // Euler walk
  std::memcpy(t, s, k - 1);
  std::vector<int> used(V.size(), 0);
  int u = 0, pos = k - 1;
  while (used[u] < (int)V[u].out.size()) {
    int v = V[u].out[used[u]++];
    t[pos++] = s[V[v].first_pos + k - 2];
    u = v;
  }
}

}  // namespace

extern "C" int rp_zscore_shuffles(const char* s1, int n1, const char* s2, int n2, int mode, unsigned int seed,
                                  int num, int k, char* out1, char* out2) {
  if (!s1 || !s2 || !out1 || !out2 || n1 < 0 || n2 < 0 || num < 0) return RP_ERR_ARG;
  if (mode != 1 && mode != 2 && mode != 12) return RP_ERR_ARG;
  GlibcRandom rng(seed);
  for (int r = 0; r < num; r++) {
    char* t1 = out1 + (size_t)r * n1;
    char* t2 = out2 + (size_t)r * n2;
    // the reference's work strings start as copies of the originals and are
    // only overwritten by the shuffles the mode enables (src/ractip.cpp:1628-1643)
    if (mode == 1 || mode == 12) klet_shuffle(s1, t1, n1, k, rng);
    else std::memcpy(t1, s1, n1);
    if (mode == 2 || mode == 12) klet_shuffle(s2, t2, n2, k, rng);
    else std::memcpy(t2, s2, n2);
  }
  return RP_OK;
}
