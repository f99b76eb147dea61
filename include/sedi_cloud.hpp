// sedi_cloud.hpp -- host-side (C++, header-only) mirror of the public interface of sediFoam's `enhancedCloud`
// (/root/reference/lammpsFoam/enhancedCloud.H:183-249) on top of the C-ABI of libsedi_b200.so.
//
// lammpsFoam.C only ever calls: the constructor (createParticles.H:7-20), evolve() (moveParticles.H:3),
// calcTcFields(), Omega(), Asrc() (liftDragCoeffs.H:16-18, pEqn.H:22) and the timers.  This class offers the same
// calls; cell fields cross the boundary as the raw storage of OpenFOAM's internal fields (Field<vector> is
// interleaved xyz doubles, Field<scalar> doubles), so a maintainer's enhancedCloud keeps its volFields and forwards
// `field.internalField().cdata()` pointers.  Per-particle data never leave the GPU.
//
// Differences to the reference that a drop-in must know (all documented in DESIGN.md):
//   * the mesh is a single-block uniform blockMesh box (every shipped case); cell = i + nx (j + ny k);
//   * diffusion smoothing (smoothField, enhancedCloud.C:790-907) runs on the GPU inside the sedi_* calls when
//     diffusionBandWidth > 0 (Jacobi-PCG on the box mesh; only the diagonal of smoothDirection acts);
//   * Omega() is identically zero, as in the reference (enhancedCloud.C:391).
#ifndef SEDI_CLOUD_HPP
#define SEDI_CLOUD_HPP
#include <string>
#include <vector>
#include "sedi_b200.h"

namespace sedi {

struct CloudProperties {          // constant/cloudProperties + transportProperties (enhancedCloud.C:573-608)
  std::string dragModel;          // "ErgunWenYu" | "SyamlalOBrien"
  int subCycles;
  double g[3];
  bool particleDrag, particlePressureGrad, particleBuoyancy, particleAddedMass, particleLift;
  bool particleHistoryForce, lubricationForce;        // enhancedCloud.C:595-598
  double inletForce[3];           // enhancedCloud.C:600-608; active inside inletBox when |inletForce| > 0
  double inletBox[9];             // x1 x2 y1 y2 z1 z2 r1 r2 - (softParticleCloud.C:471, pointInRegion :1354-1415)
  int addParticleOption;          // 1 box, 2 hollow cylinder
  double addParticleBoxEccentricity[3];
  double nub, rhob;
  double diffusionBandWidth;      // 0 = pure PCM, no smoothing (cases/.../expWachem_PCM)
  int diffusionSteps;
  double smoothDirection[3];      // diagonal of the tensor (enhancedCloud.C:578-583)
  bool UfSmooth, UpSmooth, dragSmooth, alphaSmooth;   // defaults true (enhancedCloud.C:573-576)
  CloudProperties()
      : dragModel("ErgunWenYu"), subCycles(1), particleDrag(true), particlePressureGrad(true), particleBuoyancy(false),
        particleAddedMass(false), particleLift(false), particleHistoryForce(false), lubricationForce(false), addParticleOption(0), nub(1e-6), rhob(1000.0), diffusionBandWidth(0.0), diffusionSteps(0),
        UfSmooth(true), UpSmooth(true), dragSmooth(true), alphaSmooth(true) {
    g[0] = g[1] = g[2] = 0.0; smoothDirection[0] = smoothDirection[1] = smoothDirection[2] = 1.0;
    for (int k = 0; k < 3; k++) inletForce[k] = addParticleBoxEccentricity[k] = 0.0;
    for (int k = 0; k < 9; k++) inletBox[k] = 0.0;
  }
};

class enhancedCloud {
 public:
  // lmp: the LAMMPS* / void* handle already fed with in.lammps; mesh: blockMesh box; deltaT: fluid time step
  enhancedCloud(void *lmp, const double lo[3], const double hi[3], const int ncell[3], const CloudProperties &cp, double deltaT)
      : lmp_(lmp), cp_(cp), deltaT_(deltaT) {
    sedi_mesh_box(lmp_, lo, hi, ncell);
    init();
  }
  // graded / stacked-block blockMesh: face coordinates per axis (ncell[d] + 1 values, e.g. from mesh.points()) and the
  // solver's label of every tensor cell (NULL = i + nx (j + ny k)); cases/example-cases/BL24-TH1, transport-bedload
  enhancedCloud(void *lmp, const int ncell[3], const double *xfaces, const double *yfaces, const double *zfaces, const int *cellLabel,
                const CloudProperties &cp, double deltaT)
      : lmp_(lmp), cp_(cp), deltaT_(deltaT) {
    sedi_mesh_rectilinear(lmp_, ncell, xfaces, yfaces, zfaces, cellLabel);
    init();
  }

 private:
  void init() {
    const CloudProperties &cp = cp_;
    const double deltaT = deltaT_;
    nCells_ = sedi_mesh_ncells(lmp_);
    const int model = (cp.dragModel == "SyamlalOBrien") ? SEDI_DRAG_SYAMLAL_OBRIEN_ID : SEDI_DRAG_ERGUN_WENYU_ID;
    int flags = 0;
    if (cp.particleDrag) flags |= SEDI_FORCE_DRAG_BIT;
    if (cp.particlePressureGrad) flags |= SEDI_FORCE_PGRAD_BIT;
    if (cp.particleBuoyancy) flags |= SEDI_FORCE_BUOY_BIT;
    if (cp.particleAddedMass) flags |= SEDI_FORCE_ADDEDMASS_BIT;
    if (cp.particleLift) flags |= SEDI_FORCE_LIFT_BIT;
    if (cp.particleHistoryForce) flags |= SEDI_FORCE_HISTORY_BIT;
    if (cp.lubricationForce) flags |= SEDI_FORCE_WALL_LUB_BIT;
    if (cp.addParticleOption > 0 && (cp.inletForce[0] != 0.0 || cp.inletForce[1] != 0.0 || cp.inletForce[2] != 0.0)) {
      flags |= SEDI_FORCE_INLET_BIT;
      sedi_coupling_inlet(lmp_, cp.inletForce, cp.inletBox, cp.addParticleOption, cp.addParticleBoxEccentricity);
    }
    sedi_coupling_config(lmp_, model, flags, cp.nub, cp.rhob, cp.g, deltaT);
    int sflags = 0;
    if (cp.UfSmooth) sflags |= SEDI_SMOOTH_UF_BIT;
    if (cp.UpSmooth) sflags |= SEDI_SMOOTH_UP_BIT;
    if (cp.dragSmooth) sflags |= SEDI_SMOOTH_DRAG_BIT;
    if (cp.alphaSmooth) sflags |= SEDI_SMOOTH_ALPHA_BIT;
    sedi_smooth_config(lmp_, cp.diffusionBandWidth, cp.diffusionSteps, cp.smoothDirection, sflags);
    gamma_.assign(nCells_, 0.0); Ue_.assign(3 * (size_t)nCells_, 0.0);
    Asrc_.assign(3 * (size_t)nCells_, 0.0); Omega_.assign(nCells_, 0.0);
    // softParticleCloud::adjustLampTimestep (softParticleCloud.C:209-261): dtDEM := dtFluid / round(dtFluid / dtDEM)
    const double dtIn = lammps_get_timestep(lmp_);
    long nDEM = (long)(deltaT / dtIn + 0.5);
    if (nDEM < 1) nDEM = 1;
    lammps_set_timestep(lmp_, deltaT / nDEM);
    subSteps_ = (int)(nDEM / (cp.subCycles > 0 ? cp.subCycles : 1));
    if (subSteps_ < 1) subSteps_ = 1;
    lammps_step(lmp_, 0);                      // softParticleCloud.C:189
    sedi_scatter_alpha_u(lmp_, gamma_.data(), Ue_.data());  // enhancedCloud.C:635 particleToEulerianField()
  }

 public:

  // fluid fields of the current time step (Ub, grad p, DDtUb, curl Ub): pointers to [C][3] doubles, NULL = absent
  void setFluidFields(const double *Ub, const double *gradp, const double *DDtUb, const double *curlUb) {
    sedi_put_cell_fields(lmp_, Ub, 0, gradp, DDtUb, curlUb);
    sedi_smooth_uf(lmp_);   // UfSmoothed_ = Uf (1-gamma) -> smooth -> / (1-gamma)  (enhancedCloud.C:675-690); no-op when off
  }

  // enhancedCloud::evolve(), enhancedCloud.C:669-787
  void evolve() {
    for (int k = 0; k < cp_.subCycles; k++) {
      sedi_compute_fluid_force(lmp_);          // updateParticleUr + updateDragOnParticles (:83-312)
      sedi_step(lmp_, subSteps_);              // lammpsEvolveForward (:735-743), device resident
      sedi_locate(lmp_);                       // setPositionVeloCpuId + Cloud::move (:745-753)
      if (k == 0) sedi_scatter_alpha_u(lmp_, gamma_.data(), Ue_.data());  // particleToEulerianField (:773-776)
    }
  }

  // enhancedCloud::calcTcFields(), enhancedCloud.C:316-441
  void calcTcFields() { sedi_calc_tc(lmp_, Asrc_.data(), Omega_.data()); }

  const std::vector<double> &Asrc() const { return Asrc_; }    // [C][3], kg m^-2 s^-2, consumed at pEqn.H:22
  const std::vector<double> &Omega() const { return Omega_; }  // [C], identically zero (enhancedCloud.C:391)
  const std::vector<double> &gamma() const { return gamma_; }  // [C] solid volume fraction (alpha in alphaEqn.H)
  const std::vector<double> &Ue() const { return Ue_; }        // [C][3] solid velocity (Ua)
  // ---- the rest of the public surface lammpsFoam.C / writeCPUTime.H use (enhancedCloud.H:206-249, softParticleCloud.H:351-354)
  int particleCount() const { return lammps_get_global_n(lmp_); }
  // enhancedCloud::averageInfo() prints these three lines (enhancedCloud.C:1367-1369); returned instead of printed
  struct AverageInfo { double totalVolume, totalVel[3], averageVel[3]; };
  AverageInfo averageInfo() const { AverageInfo a; sedi_average_info(lmp_, &a.totalVolume, a.totalVel, a.averageVel); return a; }
  std::vector<double> diffusionTimeCount() const { std::vector<double> t(2); sedi_get_timers(lmp_, t.data(), 0, 0); return t; }
  double particleMoveTime() const { double t = 0.0; sedi_get_timers(lmp_, 0, &t, 0); return t; }
  std::vector<double> cpuTimeSplit() const { std::vector<double> t(6); sedi_get_timers(lmp_, 0, 0, t.data()); return t; }
  // "total F before / after" and "total U solid before / after" (enhancedCloud.C:434-435, 975-976): call
  // enableConservationSums(true) once, then read them after calcTcFields() / evolve()
  void enableConservationSums(bool on) { sedi_enable_conservation_sums(lmp_, on ? 1 : 0); }
  struct ConservationSums { double Fbefore[3], Fafter[3], Ubefore[3], Uafter[3]; };
  ConservationSums conservationSums() const { ConservationSums c; sedi_get_conservation_sums(lmp_, c.Fbefore, c.Fafter, c.Ubefore, c.Uafter); return c; }
  // runTime.timeIndex() of the fluid step that is about to call evolve() (particleHistoryForce only)
  void setTimeIndex(int timeIndex) { sedi_coupling_time_index(lmp_, timeIndex); }
  int nCells() const { return nCells_; }
  int subSteps() const { return subSteps_; }
  int size() const { return lammps_get_global_n(lmp_); }

 private:
  void *lmp_;
  CloudProperties cp_;
  double deltaT_;
  int nCells_, subSteps_;
  std::vector<double> gamma_, Ue_, Asrc_, Omega_;
};

}  // namespace sedi
#endif
