// oracle/stubs: MathConst::MY_PI. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_MATH_CONST_H
#define SEDI_STUB_MATH_CONST_H
namespace LAMMPS_NS { namespace MathConst {
static const double MY_PI = 3.14159265358979323846;
static const double MY_2PI = 6.28318530717958647692;
} }
#endif
