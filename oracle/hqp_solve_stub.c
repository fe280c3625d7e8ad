/* TEST INFRASTRUCTURE ONLY.
 * The reference generates `char *hqp_solve` from hqp/hqp_solve.tcl with
 * hqp/tpc.c and Tcl_Eval's it at hqp/Hqp_Init.C:209-210.  Without a Tcl
 * interpreter the script cannot run; the outer SQP loop is restated in C++ in
 * ref_harness.cpp (hqp_solve_loop) and registered as the command "hqp_solve".
 */
char *hqp_solve = "";
