/* rkfd_ztk.cpp - ZTK reader (host only).  Grammar as visible in the reference's model files:
 * `[tag]` opens a section, `key : v v v` fields whose values may continue over following lines
 * (numbers, words, braces/commas are separators), `%` starts a comment. */
#include "rkfd_ztk.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace rkfd {
namespace {

struct Field { std::string key; std::vector<std::string> val; };
struct Section { std::string tag; std::vector<Field> fields; };

bool is_key_line(const std::string &line, std::string &key, std::string &rest)
{
  size_t i = 0; while( i < line.size() && std::isspace((unsigned char)line[i]) ) i++;
  size_t j = i; if( j >= line.size() || !(std::isalpha((unsigned char)line[j]) || line[j]=='_') ) return false;
  while( j < line.size() && (std::isalnum((unsigned char)line[j]) || line[j]=='_' || line[j]==':' || line[j]=='#') ){
    if( line[j]==':' && !( j+1 < line.size() && line[j+1]==':' ) && !( j > 0 && line[j-1]==':' ) ) break;
    j++;
  }
  size_t k = j; while( k < line.size() && std::isspace((unsigned char)line[k]) ) k++;
  if( k >= line.size() || line[k] != ':' ) return false;
  key = line.substr(i, j-i); rest = line.substr(k+1); return true;
}
void tokenize(const std::string &s, std::vector<std::string> &out)
{
  std::string t;
  for(char ch : s){
    if( std::isspace((unsigned char)ch) || ch==',' || ch=='{' || ch=='}' || ch=='(' || ch==')' || ch==';' ){ if( !t.empty() ){ out.push_back(t); t.clear(); } }
    else t.push_back(ch);
  }
  if( !t.empty() ) out.push_back(t);
}
bool parse_file(const char *fn, std::vector<Section> &secs, std::string &err)
{
  FILE *fp = std::fopen(fn, "r");
  if( !fp ){ err = std::string("cannot open ") + fn; return false; }
  char buf[4096]; Section *cur = nullptr;
  while( std::fgets(buf, sizeof buf, fp) ){
    std::string line(buf); size_t c = line.find('%'); if( c != std::string::npos ) line.erase(c);
    size_t i = 0; while( i < line.size() && std::isspace((unsigned char)line[i]) ) i++;
    if( i >= line.size() ) continue;
    if( line[i] == '[' ){ size_t e = line.find(']', i); if( e == std::string::npos ) continue;
      secs.push_back(Section()); cur = &secs.back(); cur->tag = line.substr(i+1, e-i-1); continue; }
    if( !cur ) continue;
    std::string key, rest;
    if( is_key_line(line, key, rest) ){ cur->fields.push_back(Field()); cur->fields.back().key = key; tokenize(rest, cur->fields.back().val); }
    else if( !cur->fields.empty() ) tokenize(line, cur->fields.back().val);
  }
  std::fclose(fp); return true;
}
double num(const Field &f, size_t i, double def = 0.0){ return i < f.val.size() ? std::atof(f.val[i].c_str()) : def; }
const Field *find(const Section &s, const char *key){ for(const Field &f : s.fields) if( f.key == key ) return &f; return nullptr; }
std::string word(const Section &s, const char *key){ const Field *f = find(s, key); return ( f && !f->val.empty() ) ? f->val[0] : std::string(); }
const double DEG = M_PI/180.0;

struct Shape { std::vector<double> verts; bool is_box = false; BoxShape box; };

void ring(std::vector<double> &v, const double c[3], const double ax[3], double r, int div)
{
  /* orthonormal basis perpendicular to ax */
  double a[3] = {ax[0],ax[1],ax[2]}, n = std::sqrt(a[0]*a[0]+a[1]*a[1]+a[2]*a[2]); if( n == 0 ){ a[2] = 1; n = 1; }
  for(int k=0;k<3;k++) a[k] /= n;
  double t[3] = {1,0,0}; if( std::fabs(a[0]) > 0.9 ){ t[0] = 0; t[1] = 1; }
  double d = t[0]*a[0]+t[1]*a[1]+t[2]*a[2]; double e1[3] = {t[0]-d*a[0], t[1]-d*a[1], t[2]-d*a[2]};
  n = std::sqrt(e1[0]*e1[0]+e1[1]*e1[1]+e1[2]*e1[2]); for(int k=0;k<3;k++) e1[k] /= n;
  double e2[3] = {a[1]*e1[2]-a[2]*e1[1], a[2]*e1[0]-a[0]*e1[2], a[0]*e1[1]-a[1]*e1[0]};
  for(int i=0;i<div;i++){ const double th = 2.0*M_PI*i/div, cs = std::cos(th), sn = std::sin(th);
    for(int k=0;k<3;k++) v.push_back(c[k] + r*(cs*e1[k] + sn*e2[k])); }
}

bool read_shape(const Section &s, Shape &sh, std::string &warn)
{
  const std::string type = word(s, "type");
  if( type == "box" ){
    BoxShape b; b.center[0] = b.center[1] = b.center[2] = 0; b.depth = b.width = b.height = 0;
    if( const Field *f = find(s, "center") ) for(int k=0;k<3;k++) b.center[k] = num(*f, k);
    if( const Field *f = find(s, "depth") ) b.depth = num(*f, 0);
    if( const Field *f = find(s, "width") ) b.width = num(*f, 0);
    if( const Field *f = find(s, "height") ) b.height = num(*f, 0);
    if( find(s, "ax") || find(s, "ay") || find(s, "az") ) warn += "box axes (ax/ay/az) ignored; ";
    sh.is_box = true; sh.box = b;
    for(int k=0;k<8;k++){ sh.verts.push_back(b.center[0] + ((k&1)?0.5:-0.5)*b.depth); sh.verts.push_back(b.center[1] + ((k&2)?0.5:-0.5)*b.width); sh.verts.push_back(b.center[2] + ((k&4)?0.5:-0.5)*b.height); }
    return true;
  }
  if( type == "polyhedron" ){
    if( find(s, "loop") || find(s, "prism") || find(s, "pyramid") ){ warn += "polyhedron loop/prism sugar not supported, shape '" + word(s, "name") + "' has no collision vertices; "; return true; }
    for(const Field &f : s.fields) if( f.key == "vert" && f.val.size() >= 4 ){ sh.verts.push_back(num(f,1)); sh.verts.push_back(num(f,2)); sh.verts.push_back(num(f,3)); }
    return true;
  }
  if( type == "cylinder" || type == "cone" ){
    double c[2][3] = {{0,0,0},{0,0,0}}; int nc = 0; double vert[3] = {0,0,0}; bool hasv = false;
    for(const Field &f : s.fields){
      if( f.key == "center" && nc < 2 ){ for(int k=0;k<3;k++) c[nc][k] = num(f, k); nc++; }
      if( f.key == "vert" ){ for(int k=0;k<3;k++) vert[k] = num(f, k); hasv = true; }
    }
    const Field *fr = find(s, "radius"); const double r = fr ? num(*fr, 0) : 0.0;
    const Field *fd = find(s, "div"); const int div = fd ? (int)num(*fd, 0) : 32;
    if( type == "cylinder" ){ const double ax[3] = {c[1][0]-c[0][0], c[1][1]-c[0][1], c[1][2]-c[0][2]}; ring(sh.verts, c[0], ax, r, div); ring(sh.verts, c[1], ax, r, div); }
    else { const double ax[3] = {vert[0]-c[0][0], vert[1]-c[0][1], vert[2]-c[0][2]}; if( hasv ){ sh.verts.push_back(vert[0]); sh.verts.push_back(vert[1]); sh.verts.push_back(vert[2]); } ring(sh.verts, c[0], ax, r, div); }
    return true;
  }
  warn += "shape type '" + type + "' not supported; ";
  return true;
}

}  // namespace

bool ztk_read_chain(const char *filename, ChainHost &chain, std::string &err)
{
  std::vector<Section> secs;
  if( !parse_file(filename, secs, err) ) return false;
  std::map<std::string, MotorHost> motors; std::map<std::string, Shape> shapes; std::map<std::string, int> link_index;
  std::string warn; bool has_chain = false;
  for(const Section &s : secs){
    if( s.tag == "roki::chain" || s.tag == "chain" ){ chain.name = word(s, "name"); has_chain = true; }
    else if( s.tag == "roki::motor" ){
      MotorHost m; const std::string type = word(s, "type");
      m.type = type == "dc" ? M_DC : ( type == "trq" ? M_TRQ : M_NONE );
      if( const Field *f = find(s, "motorconstant") ) m.k = num(*f, 0);
      if( const Field *f = find(s, "admittance") ) m.admittance = num(*f, 0);
      if( const Field *f = find(s, "gearratio") ) m.gear = num(*f, 0);
      if( const Field *f = find(s, "rotorinertia") ) m.rotor_inertia = num(*f, 0);
      if( const Field *f = find(s, "gearinertia") ) m.gear_inertia = num(*f, 0);
      if( const Field *f = find(s, m.type == M_DC ? "minvoltage" : "min") ) m.min = num(*f, 0);
      if( const Field *f = find(s, m.type == M_DC ? "maxvoltage" : "max") ) m.max = num(*f, 0);
      motors[word(s, "name")] = m;
    }
    else if( s.tag == "zeo::shape" ){ Shape sh; read_shape(s, sh, warn); shapes[word(s, "name")] = sh; }
    else if( s.tag == "roki::link" ){
      LinkHost l; l.name = word(s, "name"); l.stuff = word(s, "stuff");
      const std::string jt = word(s, "jointtype");
      if( jt == "fixed" || jt.empty() ) l.jtype = J_FIXED; else if( jt == "revolute" ) l.jtype = J_REVOL; else if( jt == "prismatic" ) l.jtype = J_PRISM;
      else if( jt == "spherical" ) l.jtype = J_SPHER; else if( jt == "float" ) l.jtype = J_FLOAT;
      else { err = "joint type '" + jt + "' of link '" + l.name + "' is not supported"; return false; }
      if( const Field *f = find(s, "mass") ) l.mass = num(*f, 0);
      if( const Field *f = find(s, "COM") ){ if( !f->val.empty() && f->val[0] == "auto" ) warn += "COM: auto not supported; "; else for(int k=0;k<3;k++) l.com[k] = num(*f, k); }
      if( const Field *f = find(s, "inertia") ){ if( !f->val.empty() && f->val[0] == "auto" ) warn += "inertia: auto not supported; "; else for(int k=0;k<9;k++) l.inertia[k] = num(*f, k); }
      if( const Field *f = find(s, "frame") ) for(int r=0;r<3;r++){ for(int c=0;c<3;c++) l.Ro[3*r+c] = num(*f, 4*r+c, r==c); l.po[r] = num(*f, 4*r+3); }
      if( const Field *f = find(s, "pos") ) for(int k=0;k<3;k++) l.po[k] = num(*f, k);
      if( const Field *f = find(s, "DH") ){
        /* modified DH (a, alpha, d, theta), angles in degrees ([EXT] zFrame3DFromDH) */
        const double a = num(*f,0), al = num(*f,1)*DEG, d = num(*f,2), th = num(*f,3)*DEG;
        const double sa = std::sin(al), ca = std::cos(al), st = std::sin(th), ct = std::cos(th);
        const double R[9] = { ct, -st, 0,  ca*st, ca*ct, -sa,  sa*st, sa*ct, ca };
        std::memcpy(l.Ro, R, sizeof R); l.po[0] = a; l.po[1] = -d*sa; l.po[2] = d*ca;
      }
      if( const Field *f = find(s, "stiffness") ) l.stiffness = num(*f, 0);
      if( const Field *f = find(s, "viscosity") ) l.viscosity = num(*f, 0);
      if( const Field *f = find(s, "coulomb") ) l.coulomb = num(*f, 0);
      if( const Field *f = find(s, "staticfriction") ) l.sfriction = num(*f, 0);
      const std::string mn = word(s, "motor");
      if( !mn.empty() ){ auto it = motors.find(mn); if( it == motors.end() ){ err = "unknown motor '" + mn + "'"; return false; } l.motor = it->second; }
      const std::string pn = word(s, "parent");
      if( !pn.empty() ){ auto it = link_index.find(pn); if( it == link_index.end() ){ err = "parent '" + pn + "' of link '" + l.name + "' must be defined before it"; return false; } l.parent = it->second; }
      for(const Field &f : s.fields) if( f.key == "shape" && !f.val.empty() ){
        auto it = shapes.find(f.val[0]); if( it == shapes.end() ){ err = "unknown shape '" + f.val[0] + "'"; return false; }
        if( !it->second.verts.empty() ) l.shapes.push_back(it->second.verts);
        if( it->second.is_box ) l.boxes.push_back(it->second.box);
      }
      link_index[l.name] = (int)chain.links.size();
      chain.links.push_back(l);
    }
  }
  if( !has_chain && chain.links.empty() ){ err = std::string("no [roki::chain] in ") + filename; return false; }
  chain.sync_sizes();
  for(const Section &s : secs) if( s.tag == "roki::chain::init" ){
    if( const Field *f = find(s, "frame") ) if( !chain.links.empty() ){
      /* root frame override: composed onto the root link's org frame */
      LinkHost &r = chain.links[0]; double R[9], p[3];
      for(int a=0;a<3;a++){ for(int c=0;c<3;c++) R[3*a+c] = num(*f, 4*a+c, a==c); p[a] = num(*f, 4*a+3); }
      double Rn[9], pn[3];
      for(int a=0;a<3;a++){ for(int c=0;c<3;c++){ Rn[3*a+c] = 0; for(int k=0;k<3;k++) Rn[3*a+c] += R[3*a+k]*r.Ro[3*k+c]; }
        pn[a] = p[a]; for(int k=0;k<3;k++) pn[a] += R[3*a+k]*r.po[k]; }
      std::memcpy(r.Ro, Rn, sizeof Rn); std::memcpy(r.po, pn, sizeof pn);
    }
    for(const Field &f : s.fields) if( f.key == "joint" && !f.val.empty() ){
      auto it = link_index.find(f.val[0]); if( it == link_index.end() ) continue;
      const LinkHost &l = chain.links[it->second]; const int o = chain.link_qofs(it->second), n = jtype_ndof(l.jtype);
      for(int k=0;k<n;k++){
        double v = num(f, 1+k);
        const bool angular = l.jtype == J_REVOL || l.jtype == J_SPHER || ( l.jtype == J_FLOAT && k >= 3 );
        chain.dis[o+k] = angular ? v*DEG : v;      /* [EXT] angles are written in degrees */
      }
    }
  }
  if( !warn.empty() ) std::fprintf(stderr, "rokifd_b200: %s: %s\n", filename, warn.c_str());
  return true;
}

bool ztk_read_contact_info(const char *filename, std::vector<ContactInfoHost> &ci, std::string &err)
{
  std::vector<Section> secs;
  if( !parse_file(filename, secs, err) ) return false;
  for(const Section &s : secs){
    if( s.tag != "roki::contact" && s.tag != "contact" ) continue;
    ContactInfoHost c; const Field *b = find(s, "bind");
    if( !b || b->val.size() < 2 ){ err = "[roki::contact] without `bind: a b`"; return false; }
    c.a = b->val[0]; c.b = b->val[1];
    if( const Field *f = find(s, "staticfriction") ) c.SF = num(*f, 0);
    if( const Field *f = find(s, "kineticfriction") ) c.KF = num(*f, 0);
    /* a record with elasticity/viscosity is ELASTIC, with compensation/relaxation RIGID ([EXT A-11]) */
    if( find(s, "elasticity") || find(s, "viscosity") ){
      c.type = C_ELASTIC;
      if( const Field *f = find(s, "elasticity") ) c.E = num(*f, 0);
      if( const Field *f = find(s, "viscosity") ) c.V = num(*f, 0);
    } else {
      c.type = C_RIGID;
      if( const Field *f = find(s, "compensation") ) c.K = num(*f, 0);
      if( const Field *f = find(s, "relaxation") ) c.L = num(*f, 0);
    }
    ci.push_back(c);
  }
  return true;
}

}  // namespace rkfd
