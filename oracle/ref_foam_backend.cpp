// oracle/ref_foam_backend.cpp -- TEST INFRASTRUCTURE ONLY.  Built only where /root/reference exists, into
// oracle/_ref/libsedi_ref.so.  Instantiates the reference's OWN drag closures -- lammpsFoam/dragModels/
// {dragModel/dragModel.C, ErgunWenYu/ErgunWenYu.C, SyamlalOBrien/SyamlalOBrien.C}, compiled unmodified and by path
// against oracle/stubs_foam/ -- and evaluates Jd on caller arrays.  Nothing of the reference is copied here.
#include "ErgunWenYu.H"
#include "SyamlalOBrien.H"

namespace {
template <class Model>
void eval_jd(int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof, double *out) {
  Foam::dictionary cloudDict;
  Foam::IOdictionary transDict;
  transDict.set("nub", nuf);
  transDict.set("rhob", rhof);
  const Foam::scalarField a(alpha, n), d(pd, n), u(Ur, n);
  Model model(cloudDict, transDict, a, d);          // ErgunWenYu.C:48-74 / SyamlalOBrien.C:48-74
  Foam::tmp<Foam::scalarField> jd = model.Jd(u);    // ErgunWenYu.C:86-145 / SyamlalOBrien.C:85-144
  for (int i = 0; i < n; i++) out[i] = jd()[i];
}
}  // namespace

extern "C" {
void ora_ref_jd_ergun_wenyu(int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof, double *out) {
  eval_jd<Foam::ErgunWenYu>(n, Ur, alpha, pd, nuf, rhof, out);
}
void ora_ref_jd_syamlal_obrien(int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof, double *out) {
  eval_jd<Foam::SyamlalOBrien>(n, Ur, alpha, pd, nuf, rhof, out);
}
}
