/*
 * Prg_DIDCuda.C -- the reference's own example program (hqp_docp/Prg_DID.C: double
 * integrator with a state constraint) with its stage loop on the GPU: SURVEY.md section 8,
 * row f4.  Everything of Prg_DID is inherited (horizon, variables, bounds, structure,
 * update_vals for prg_simulate); Hqp_DocpCuda<> replaces Hqp_Docp::update / ::update_fbd by
 * one device call each, evaluating ModelDID (hqp_b200/csrc/docp_models.cuh), the device
 * counterpart of Prg_DID::update_vals (hqp_docp/Prg_DID.C:78-98).
 * Selectable like any program of the reference: "prg_name DIDCuda".
 */
#include "Hqp_DocpCuda.h"

#include <Prg_DID.h>
#include <If_Class.h>

class Prg_DIDCuda : public Hqp_DocpCuda<Prg_DID> {
 public:
  const char *name() { return "DIDCuda"; }
  int cuda_model() { return HQPDOCP_MODEL_DID; }
  void cuda_params(int K, std::vector<double> &par, int &nspar, std::vector<double> &) {
    par.assign(1, 1.0 / K);  // dt = 1.0 / _kmax, hqp_docp/Prg_DID.C:82
    nspar = 0;
  }
};

IF_CLASS_DEFINE("DIDCuda", Prg_DIDCuda, Hqp_SqpProgram);
