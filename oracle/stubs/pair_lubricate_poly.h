// oracle/stubs: stock PairLubricate state + the PairLubricatePoly declaration. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_PAIR_LUBRICATE_POLY_H
#define SEDI_STUB_PAIR_LUBRICATE_POLY_H
#include "pair.h"
namespace LAMMPS_NS {
class FixWall;
class PairLubricate : public Pair {
 public:
  double mu, cut_inner_global, cut_global; int flaglog, flagfld, shearing, flagHI, flagVF, flagdeform, flagwall;
  double vol_P; FixWall *wallfix; double Ef[3][3]; double R0, RT0, RS0; double **cut_inner, **cut;
  PairLubricate(LAMMPS *l) : Pair(l), mu(0), cut_inner_global(0), cut_global(0), flaglog(0), flagfld(0),
    shearing(0), flagHI(1), flagVF(1), flagdeform(0), flagwall(0), vol_P(0), wallfix(0), R0(0), RT0(0), RS0(0),
    cut_inner(0), cut(0) {}
  virtual void compute(int, int) {}
  virtual void settings(int, char **) {}
};
class PairLubricatePoly : public PairLubricate {
 public:
  PairLubricatePoly(LAMMPS *);
  void compute(int, int);
  void init_style();
};
}
#endif
