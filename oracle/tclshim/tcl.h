/*
 * Declarations-only stand-in for <tcl.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  Tcl is not installed in this image; the
 * reference (omuses/hqp) registers every option and command through
 * iftcl -> Tcl_CreateObjCommand.  This header declares exactly the subset of
 * the Tcl C API that the reference objects reference (iftcl/*.C, hqp/*.C,
 * iftcl/If.h:76-85), so that the UNMODIFIED reference sources compile where
 * they lie under /root/reference.  The implementations are in tclshim.cpp.
 * Nothing here is Tcl source; it is a minimal command table.
 */
#ifndef HQP_ORACLE_TCL_SHIM_H
#define HQP_ORACLE_TCL_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

#define TCL_OK 0
#define TCL_ERROR 1
#define TCL_GLOBAL_ONLY 1
#define TCL_VERSION "8.6"
#define TCL_MAJOR_VERSION 8
#define TCL_MINOR_VERSION 6
#ifndef CONST
#define CONST const
#endif
#ifndef CONST84
#define CONST84 const
#endif

typedef struct Tcl_Interp Tcl_Interp;
typedef struct Tcl_Obj Tcl_Obj;
typedef void *ClientData;
typedef struct Tcl_Command_ *Tcl_Command;
typedef int(Tcl_ObjCmdProc)(ClientData clientData, Tcl_Interp *interp, int objc,
                            Tcl_Obj *CONST objv[]);
typedef void(Tcl_CmdDeleteProc)(ClientData clientData);
typedef int(Tcl_AppInitProc)(Tcl_Interp *interp);

Tcl_Interp *Tcl_CreateInterp(void);
int Tcl_Init(Tcl_Interp *interp);
void Tcl_FindExecutable(const char *argv0);
const char *Tcl_InitStubs(Tcl_Interp *interp, const char *version, int exact);
int Tcl_PkgProvide(Tcl_Interp *interp, const char *name, const char *version);
void Tcl_Main(int argc, char **argv, Tcl_AppInitProc *appInitProc);
const char *Tcl_SetVar(Tcl_Interp *interp, const char *varName,
                       const char *newValue, int flags);

Tcl_Command Tcl_CreateObjCommand(Tcl_Interp *interp, const char *cmdName,
                                 Tcl_ObjCmdProc *proc, ClientData clientData,
                                 Tcl_CmdDeleteProc *deleteProc);
int Tcl_DeleteCommandFromToken(Tcl_Interp *interp, Tcl_Command token);
const char *Tcl_GetCommandName(Tcl_Interp *interp, Tcl_Command token);

int Tcl_Eval(Tcl_Interp *interp, const char *script);
int Tcl_VarEval(Tcl_Interp *interp, ...);
int Tcl_EvalObjv(Tcl_Interp *interp, int objc, Tcl_Obj *CONST objv[],
                 int flags);

void Tcl_AppendResult(Tcl_Interp *interp, ...);
void Tcl_ResetResult(Tcl_Interp *interp);
Tcl_Obj *Tcl_GetObjResult(Tcl_Interp *interp);
void Tcl_SetObjResult(Tcl_Interp *interp, Tcl_Obj *obj);
const char *Tcl_GetStringResult(Tcl_Interp *interp);

Tcl_Obj *Tcl_NewStringObj(const char *bytes, int length);
Tcl_Obj *Tcl_NewIntObj(int v);
Tcl_Obj *Tcl_NewDoubleObj(double v);
Tcl_Obj *Tcl_NewBooleanObj(int v);
Tcl_Obj *Tcl_NewListObj(int objc, Tcl_Obj *CONST objv[]);
int Tcl_ListObjAppendElement(Tcl_Interp *interp, Tcl_Obj *list, Tcl_Obj *obj);
int Tcl_ListObjGetElements(Tcl_Interp *interp, Tcl_Obj *list, int *objc,
                           Tcl_Obj ***objv);
char *Tcl_GetString(Tcl_Obj *obj);
char *Tcl_GetStringFromObj(Tcl_Obj *obj, int *length);
int Tcl_GetIntFromObj(Tcl_Interp *interp, Tcl_Obj *obj, int *out);
int Tcl_GetDoubleFromObj(Tcl_Interp *interp, Tcl_Obj *obj, double *out);
int Tcl_GetBooleanFromObj(Tcl_Interp *interp, Tcl_Obj *obj, int *out);

void TclShim_DecrRefCount(Tcl_Obj *obj);
void TclShim_IncrRefCount(Tcl_Obj *obj);
#define Tcl_DecrRefCount(o) TclShim_DecrRefCount(o)
#define Tcl_IncrRefCount(o) TclShim_IncrRefCount(o)

#ifdef __cplusplus
}
#endif

#endif
