/*
 * Minimal command-table implementation of the Tcl C API subset declared in
 * tcl.h (this directory).
 *
 * TEST INFRASTRUCTURE ONLY: lets the unmodified reference (iftcl + hqp) link
 * and run in an image without Tcl.  Semantics that the reference relies on:
 *   - re-creating a command name first runs the old command's delete proc
 *     (iftcl/If_Element.C:65-68 marks the old element deleted; e.g. "sqp_init"
 *     is registered twice, hqp/Hqp_SqpSolver.C:119 and hqp/Hqp_SqpPowell.C:68);
 *   - Tcl_Eval handles "cmd", "cmd value", "cmd {braced value}", several
 *     commands separated by newline/semicolon, and silently accepts the few
 *     core-Tcl words the reference evaluates for side effects only
 *     (puts/update/if/proc/..., iftcl/If.C:55-56, hqp/Hqp_Init.C:210-216);
 *   - numbers are rendered with %.17g so If_GetReal round-trips exactly.
 */
#include "tcl.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct Tcl_Obj {
  std::string s;
  std::vector<Tcl_Obj *> list;
  bool is_list = false;
  int refs = 0;
};

struct Tcl_Command_ {
  std::string name;
  Tcl_ObjCmdProc *proc = nullptr;
  ClientData cd = nullptr;
  Tcl_CmdDeleteProc *del = nullptr;
  bool alive = false;
};

struct Tcl_Interp {
  std::map<std::string, Tcl_Command_ *> cmds;
  Tcl_Obj *res = nullptr;
};

static Tcl_Obj *new_obj(const std::string &s) {
  Tcl_Obj *o = new Tcl_Obj;
  o->s = s;
  return o;
}

static void free_obj(Tcl_Obj *o) {
  for (Tcl_Obj *e : o->list) TclShim_DecrRefCount(e);
  delete o;
}

static void render_list(Tcl_Obj *o) {
  if (!o->is_list) return;
  std::string out;
  for (size_t i = 0; i < o->list.size(); i++) {
    Tcl_Obj *e = o->list[i];
    render_list(e);
    if (i) out += ' ';
    bool brace = e->s.empty() || e->s.find_first_of(" \t\n") != std::string::npos;
    out += brace ? "{" + e->s + "}" : e->s;
  }
  o->s = out;
}

/* split a Tcl word list honouring {..} and ".." grouping */
static bool split_words(const char *p, const char *end,
                        std::vector<std::string> &words) {
  while (p < end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    if (p >= end) break;
    std::string w;
    if (*p == '{') {
      int depth = 1;
      p++;
      while (p < end && depth > 0) {
        if (*p == '{') depth++;
        if (*p == '}') {
          depth--;
          if (depth == 0) break;
        }
        w += *p++;
      }
      if (depth != 0) return false;
      p++;
    } else if (*p == '"') {
      p++;
      while (p < end && *p != '"') w += *p++;
      if (p >= end) return false;
      p++;
    } else {
      while (p < end && *p != ' ' && *p != '\t' && *p != '\r') w += *p++;
    }
    words.push_back(w);
  }
  return true;
}

static void set_result(Tcl_Interp *ip, Tcl_Obj *o) {
  if (o) o->refs++;
  if (ip->res) TclShim_DecrRefCount(ip->res);
  ip->res = o;
}

static int eval_words(Tcl_Interp *ip, std::vector<Tcl_Obj *> &objv) {
  Tcl_ResetResult(ip);
  if (objv.empty()) return TCL_OK;
  const std::string &name = objv[0]->s;
  auto it = ip->cmds.find(name);
  if (it == ip->cmds.end() || !it->second->alive) {
    static const char *ignored[] = {"puts", "update", "if", "proc", "catch",
                                    "set",  "global", "package", "source",
                                    "namespace", "#", nullptr};
    for (const char **q = ignored; *q; q++)
      if (name == *q || name[0] == '#') return TCL_OK;
    Tcl_AppendResult(ip, "invalid command name \"", name.c_str(), "\"", NULL);
    return TCL_ERROR;
  }
  Tcl_Command_ *c = it->second;
  return c->proc(c->cd, ip, (int)objv.size(), objv.data());
}

extern "C" {

Tcl_Interp *Tcl_CreateInterp(void) { return new Tcl_Interp; }
int Tcl_Init(Tcl_Interp *) { return TCL_OK; }
void Tcl_FindExecutable(const char *) {}
const char *Tcl_InitStubs(Tcl_Interp *, const char *, int) { return TCL_VERSION; }
int Tcl_PkgProvide(Tcl_Interp *, const char *, const char *) { return TCL_OK; }
void Tcl_Main(int, char **, Tcl_AppInitProc *) {}
const char *Tcl_SetVar(Tcl_Interp *, const char *, const char *v, int) { return v; }

Tcl_Command Tcl_CreateObjCommand(Tcl_Interp *ip, const char *cmdName,
                                 Tcl_ObjCmdProc *proc, ClientData cd,
                                 Tcl_CmdDeleteProc *del) {
  auto it = ip->cmds.find(cmdName);
  if (it != ip->cmds.end() && it->second->alive) {
    Tcl_Command_ *old = it->second;
    old->alive = false;
    if (old->del) old->del(old->cd);
  }
  Tcl_Command_ *c = new Tcl_Command_; /* never freed: tokens stay valid */
  c->name = cmdName;
  c->proc = proc;
  c->cd = cd;
  c->del = del;
  c->alive = true;
  ip->cmds[cmdName] = c;
  return c;
}

int Tcl_DeleteCommandFromToken(Tcl_Interp *ip, Tcl_Command token) {
  if (!token || !token->alive) return -1;
  token->alive = false;
  auto it = ip->cmds.find(token->name);
  if (it != ip->cmds.end() && it->second == token) ip->cmds.erase(it);
  if (token->del) token->del(token->cd);
  return 0;
}

const char *Tcl_GetCommandName(Tcl_Interp *, Tcl_Command token) {
  return token ? token->name.c_str() : "";
}

int Tcl_EvalObjv(Tcl_Interp *ip, int objc, Tcl_Obj *CONST objv[], int) {
  std::vector<Tcl_Obj *> v(objv, objv + objc);
  for (Tcl_Obj *o : v) render_list(o);
  return eval_words(ip, v);
}

int Tcl_Eval(Tcl_Interp *ip, const char *script) {
  if (!script) return TCL_OK;
  const char *p = script;
  const char *end = script + strlen(script);
  int rc = TCL_OK;
  while (p < end) {
    /* one command = up to newline / semicolon outside braces */
    const char *q = p;
    int depth = 0;
    while (q < end) {
      if (*q == '{') depth++;
      if (*q == '}') depth--;
      if (depth <= 0 && (*q == '\n' || *q == ';')) break;
      q++;
    }
    std::vector<std::string> words;
    if (!split_words(p, q, words)) {
      Tcl_ResetResult(ip);
      Tcl_AppendResult(ip, "unbalanced braces or quotes", NULL);
      return TCL_ERROR;
    }
    if (!words.empty()) {
      std::vector<Tcl_Obj *> objv;
      for (auto &w : words) {
        Tcl_Obj *o = new_obj(w);
        o->refs = 1;
        objv.push_back(o);
      }
      rc = eval_words(ip, objv);
      for (Tcl_Obj *o : objv) TclShim_DecrRefCount(o);
      if (rc != TCL_OK) return rc;
    }
    p = q + 1;
  }
  return rc;
}

int Tcl_VarEval(Tcl_Interp *ip, ...) {
  std::string script;
  va_list ap;
  va_start(ap, ip);
  for (const char *s = va_arg(ap, const char *); s; s = va_arg(ap, const char *))
    script += s;
  va_end(ap);
  return Tcl_Eval(ip, script.c_str());
}

void Tcl_AppendResult(Tcl_Interp *ip, ...) {
  std::string add;
  va_list ap;
  va_start(ap, ip);
  for (const char *s = va_arg(ap, const char *); s; s = va_arg(ap, const char *))
    add += s;
  va_end(ap);
  std::string cur = ip->res ? (render_list(ip->res), ip->res->s) : std::string();
  set_result(ip, new_obj(cur + add));
}

void Tcl_ResetResult(Tcl_Interp *ip) { set_result(ip, nullptr); }

Tcl_Obj *Tcl_GetObjResult(Tcl_Interp *ip) {
  if (!ip->res) set_result(ip, new_obj(""));
  return ip->res;
}

void Tcl_SetObjResult(Tcl_Interp *ip, Tcl_Obj *obj) { set_result(ip, obj); }

const char *Tcl_GetStringResult(Tcl_Interp *ip) {
  Tcl_Obj *o = Tcl_GetObjResult(ip);
  render_list(o);
  return o->s.c_str();
}

Tcl_Obj *Tcl_NewStringObj(const char *bytes, int length) {
  if (!bytes) return new_obj("");
  return new_obj(length < 0 ? std::string(bytes) : std::string(bytes, length));
}

Tcl_Obj *Tcl_NewIntObj(int v) { return new_obj(std::to_string(v)); }

Tcl_Obj *Tcl_NewDoubleObj(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  return new_obj(buf);
}

Tcl_Obj *Tcl_NewBooleanObj(int v) { return new_obj(v ? "1" : "0"); }

Tcl_Obj *Tcl_NewListObj(int objc, Tcl_Obj *CONST objv[]) {
  Tcl_Obj *o = new_obj("");
  o->is_list = true;
  for (int i = 0; i < objc; i++) {
    objv[i]->refs++;
    o->list.push_back(objv[i]);
  }
  return o;
}

int Tcl_ListObjAppendElement(Tcl_Interp *, Tcl_Obj *list, Tcl_Obj *obj) {
  list->is_list = true;
  obj->refs++;
  list->list.push_back(obj);
  return TCL_OK;
}

int Tcl_ListObjGetElements(Tcl_Interp *ip, Tcl_Obj *o, int *objc,
                           Tcl_Obj ***objv) {
  if (!o->is_list) {
    std::vector<std::string> words;
    if (!split_words(o->s.c_str(), o->s.c_str() + o->s.size(), words)) {
      if (ip) Tcl_AppendResult(ip, "malformed list", NULL);
      return TCL_ERROR;
    }
    for (auto &w : words) {
      Tcl_Obj *e = new_obj(w);
      e->refs = 1;
      o->list.push_back(e);
    }
    o->is_list = true;
  }
  *objc = (int)o->list.size();
  *objv = o->list.data();
  return TCL_OK;
}

char *Tcl_GetString(Tcl_Obj *o) {
  render_list(o);
  return const_cast<char *>(o->s.c_str());
}

char *Tcl_GetStringFromObj(Tcl_Obj *o, int *length) {
  render_list(o);
  if (length) *length = (int)o->s.size();
  return const_cast<char *>(o->s.c_str());
}

int Tcl_GetIntFromObj(Tcl_Interp *ip, Tcl_Obj *o, int *out) {
  render_list(o);
  char *e = nullptr;
  long v = strtol(o->s.c_str(), &e, 0);
  while (e && (*e == ' ' || *e == '\t')) e++;
  if (o->s.empty() || !e || *e) {
    if (ip) {
      Tcl_ResetResult(ip);
      Tcl_AppendResult(ip, "expected integer but got \"", o->s.c_str(), "\"", NULL);
    }
    return TCL_ERROR;
  }
  *out = (int)v;
  return TCL_OK;
}

int Tcl_GetDoubleFromObj(Tcl_Interp *ip, Tcl_Obj *o, double *out) {
  render_list(o);
  char *e = nullptr;
  double v = strtod(o->s.c_str(), &e);
  while (e && (*e == ' ' || *e == '\t')) e++;
  if (o->s.empty() || !e || *e) {
    if (ip) {
      Tcl_ResetResult(ip);
      Tcl_AppendResult(ip, "expected floating-point number but got \"",
                       o->s.c_str(), "\"", NULL);
    }
    return TCL_ERROR;
  }
  *out = v;
  return TCL_OK;
}

int Tcl_GetBooleanFromObj(Tcl_Interp *ip, Tcl_Obj *o, int *out) {
  render_list(o);
  const std::string &s = o->s;
  if (s == "1" || s == "true" || s == "yes" || s == "on") { *out = 1; return TCL_OK; }
  if (s == "0" || s == "false" || s == "no" || s == "off") { *out = 0; return TCL_OK; }
  int iv;
  if (Tcl_GetIntFromObj(nullptr, o, &iv) == TCL_OK) { *out = iv != 0; return TCL_OK; }
  if (ip) {
    Tcl_ResetResult(ip);
    Tcl_AppendResult(ip, "expected boolean value but got \"", s.c_str(), "\"", NULL);
  }
  return TCL_ERROR;
}

void TclShim_IncrRefCount(Tcl_Obj *o) { o->refs++; }
void TclShim_DecrRefCount(Tcl_Obj *o) {
  if (--o->refs <= 0) free_obj(o);
}

} /* extern "C" */
