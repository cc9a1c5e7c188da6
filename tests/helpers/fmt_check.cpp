// Test helper (tests/test_host_formats.py): the Fortran edit-descriptor emulation of the command-line twins, callable from Python.
#include "../../cdftools_b200/csrc/host/host_common.hpp"

extern "C" {
void fmt_f(double x, int w, int d, char *out) { strcpy(out, cdfhost::fortran_f(x, w, d).c_str()); }
void fmt_e16_7(double x, char *out) { strcpy(out, cdfhost::fortran_e16_7(x).c_str()); }
}
