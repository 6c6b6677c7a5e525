// Constrained-minimisation (Ern-Vohralik) patch kernel - see DESIGN.md.
#include "eqlb_internal.cuh"

void launch_ev(eqlb_handle* h, const double* const* dG, const double* const* dF, double* const* dSigma)
{
  (void)h; (void)dG; (void)dF; (void)dSigma;
  throw EqlbError(EQLB_ERR_STATE, "eqlb_ev_run: EV kernel not built into this library");
}
