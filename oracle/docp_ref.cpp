// The hqp_docp example (hqp_docp/Docp_Main.C:21-110) on top of the harness.
// TEST INFRASTRUCTURE ONLY.
//   docp_ref [kmax] [qp_solver] [mat_solver] [plugin.so] [with_cns] [verbose]
#include <cstdio>
#include <cstdlib>
extern "C" {
int ref_init(void);
void ref_set_verbose(int);
int ref_load_plugin(const char *);
int ref_docp_did(int, const char *, const char *, double, int, double *, int *,
                 int *, int *, char *, int);
}
int main(int argc, char **argv) {
  int kmax = argc > 1 ? atoi(argv[1]) : 60;
  const char *qps = argc > 2 ? argv[2] : "";
  const char *mat = argc > 3 ? argv[3] : "LQDOCP";
  const char *plugin = argc > 4 ? argv[4] : "";
  int with_cns = argc > 5 ? atoi(argv[5]) : 1;
  int verbose = argc > 6 ? atoi(argv[6]) : 1;
  ref_init();
  ref_set_verbose(verbose);
  if (*plugin && ref_load_plugin(plugin)) return 2;
  double obj = 0;
  int sqp_it = 0, qp_it = 0, steps = 0;
  char res[256];
  int rc = ref_docp_did(kmax, qps, mat, 1e-5, with_cns, &obj, &sqp_it, &qp_it,
                        &steps, res, sizeof res);
  if (rc) {
    printf("{\"error\": %d}\n", rc);
    return 1;
  }
  printf("{\"result\": \"%s\", \"objective\": %.13g, \"sqp_iters\": %d, "
         "\"qp_iters\": %d, \"line_steps\": %d}\n",
         res, obj, sqp_it, qp_it, steps);
  return 0;
}
