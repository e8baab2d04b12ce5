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

/* mass properties of a shape at unit density, about the origin of the frame the shape is written in ([EXT] Zeo computes
 * them for `COM: auto` / `inertia: auto`): volume, volume * centroid, second-moment matrix int x x^T dV */
struct MassProp {
  double vol = 0, vc[3] = {0,0,0}, xx[9] = {0,0,0, 0,0,0, 0,0,0};
  void add(const MassProp &o){ vol += o.vol; for(int i=0;i<3;i++) vc[i] += o.vc[i]; for(int i=0;i<9;i++) xx[i] += o.xx[i]; }
};
struct Shape { std::vector<double> verts; bool is_box = false; BoxShape box; MassProp mp; bool has_mp = false; };

/* closed triangle mesh: signed tetrahedra against the origin */
MassProp mesh_massprop(const std::vector<double> &v, const std::vector<int> &tri)
{
  MassProp m;
  for(size_t t=0;t+2<tri.size();t+=3){
    const double *a = &v[3*tri[t]], *b = &v[3*tri[t+1]], *c = &v[3*tri[t+2]];
    const double vol = ( a[0]*(b[1]*c[2]-b[2]*c[1]) + a[1]*(b[2]*c[0]-b[0]*c[2]) + a[2]*(b[0]*c[1]-b[1]*c[0]) )/6.0;
    m.vol += vol;
    double S[3];
    for(int i=0;i<3;i++){ S[i] = a[i]+b[i]+c[i]; m.vc[i] += vol*0.25*S[i]; }
    for(int i=0;i<3;i++) for(int j=0;j<3;j++) m.xx[3*i+j] += vol/20.0*( a[i]*a[j] + b[i]*b[j] + c[i]*c[j] + S[i]*S[j] );
  }
  if( m.vol < 0 ){ m.vol = -m.vol; for(int i=0;i<3;i++) m.vc[i] = -m.vc[i]; for(int i=0;i<9;i++) m.xx[i] = -m.xx[i]; }   /* inward-oriented faces */
  return m;
}
/* a solid of revolution-like primitive given in its own axes (e1, e2, ax unit, origin o): volume V, centroid at o + zc ax,
 * second moments about the centroid diag(it, it, ia) in (e1, e2, ax) */
MassProp axial_massprop(const double o[3], const double ax[3], double V, double zc, double it, double ia)
{
  MassProp m; m.vol = V; double c[3];
  for(int i=0;i<3;i++){ c[i] = o[i] + zc*ax[i]; m.vc[i] = V*c[i]; }
  /* second moments about the centroid: it * (I - ax ax^T) + ia * ax ax^T, then the parallel shift V c c^T */
  for(int i=0;i<3;i++) for(int j=0;j<3;j++) m.xx[3*i+j] = it*((i==j?1.0:0.0) - ax[i]*ax[j]) + ia*ax[i]*ax[j] + V*c[i]*c[j];
  return m;
}

void basis_perp(const double ax[3], double a[3], double e1[3], double e2[3])
{
  double n = std::sqrt(ax[0]*ax[0]+ax[1]*ax[1]+ax[2]*ax[2]); a[0] = ax[0]; a[1] = ax[1]; a[2] = ax[2]; if( n == 0 ){ a[0] = a[1] = 0; a[2] = 1; n = 1; }
  for(int k=0;k<3;k++) a[k] /= n;
  double t[3] = {1,0,0}; if( std::fabs(a[0]) > 0.9 ){ t[0] = 0; t[1] = 1; }
  const double d = t[0]*a[0]+t[1]*a[1]+t[2]*a[2];
  for(int k=0;k<3;k++) e1[k] = t[k]-d*a[k];
  n = std::sqrt(e1[0]*e1[0]+e1[1]*e1[1]+e1[2]*e1[2]); for(int k=0;k<3;k++) e1[k] /= n;
  e2[0] = a[1]*e1[2]-a[2]*e1[1]; e2[1] = a[2]*e1[0]-a[0]*e1[2]; e2[2] = a[0]*e1[1]-a[1]*e1[0];
}
void ring(std::vector<double> &v, const double c[3], const double ax[3], double r, int div)
{
  double a[3], e1[3], e2[3]; basis_perp(ax, a, e1, e2);
  for(int i=0;i<div;i++){ const double th = 2.0*M_PI*i/div, cs = std::cos(th), sn = std::sin(th);
    for(int k=0;k<3;k++) v.push_back(c[k] + r*(cs*e1[k] + sn*e2[k])); }
}

/* `loop: <axis> <coordinate>  x y | arc cw|ccw <radius> <div> ...` + `prism: dx dy dz` | `pyramid: x y z` ([EXT] Zeo's
 * polyhedron sugar, puma.ztk:67-88): a planar polygon whose corners may be joined by circular arcs (radius r, the minor
 * arc, `div` segments), extruded along a vector or joined to an apex.  Vertices: the loop, then the shifted loop / apex. */
bool read_loop_sugar(const Section &s, Shape &sh, std::string &warn)
{
  const Field *fl = find(s, "loop"); if( !fl || fl->val.size() < 2 ) return false;
  const char axc = fl->val[0].empty() ? 'z' : fl->val[0][0]; const double h0 = std::atof(fl->val[1].c_str());
  std::vector<double> pts;      /* 2-D loop */
  struct Arc { size_t after; bool cw; double r; int div; }; std::vector<Arc> arcs;
  for(size_t i=2;i<fl->val.size();){
    if( fl->val[i] == "arc" && i+3 < fl->val.size() ){ Arc a; a.after = pts.size()/2; a.cw = fl->val[i+1] == "cw"; a.r = std::atof(fl->val[i+2].c_str()); a.div = std::atoi(fl->val[i+3].c_str()); arcs.push_back(a); i += 4; }
    else if( i+1 < fl->val.size() ){ pts.push_back(std::atof(fl->val[i].c_str())); pts.push_back(std::atof(fl->val[i+1].c_str())); i += 2; }
    else break;
  }
  const size_t np = pts.size()/2; if( np < 3 ){ warn += "polyhedron loop with fewer than 3 corners; "; return false; }
  std::vector<double> loop;
  for(size_t k=0;k<np;k++){
    loop.push_back(pts[2*k]); loop.push_back(pts[2*k+1]);
    for(const Arc &a : arcs) if( a.after == k+1 ){       /* between corner k and corner k+1 (cyclic) */
      const double *p0 = &pts[2*k], *p1 = &pts[2*((k+1)%np)];
      const double dx = p1[0]-p0[0], dy = p1[1]-p0[1], ch = std::sqrt(dx*dx+dy*dy); if( ch == 0 || a.div < 2 ) continue;
      const double r = a.r > 0.5*ch ? a.r : 0.5*ch, hh = std::sqrt(r*r - 0.25*ch*ch);
      /* clockwise travel keeps the centre on the right of the direction of travel */
      const double sg = a.cw ? 1.0 : -1.0, cx = 0.5*(p0[0]+p1[0]) + sg*hh*dy/ch, cy = 0.5*(p0[1]+p1[1]) - sg*hh*dx/ch;
      double a0 = std::atan2(p0[1]-cy, p0[0]-cx), a1 = std::atan2(p1[1]-cy, p1[0]-cx), sw = a.cw ? a0 - a1 : a1 - a0;
      while( sw <= 0 ) sw += 2.0*M_PI;
      while( sw > 2.0*M_PI ) sw -= 2.0*M_PI;
      for(int j=1;j<a.div;j++){ const double th = a0 + (a.cw ? -1.0 : 1.0)*sw*j/a.div; loop.push_back(cx + r*std::cos(th)); loop.push_back(cy + r*std::sin(th)); }
    }
  }
  const size_t n = loop.size()/2;
  auto to3 = [&](double u, double v, double out[3]){ if( axc == 'x' ){ out[0] = h0; out[1] = u; out[2] = v; } else if( axc == 'y' ){ out[0] = v; out[1] = h0; out[2] = u; } else { out[0] = u; out[1] = v; out[2] = h0; } };
  std::vector<int> tri;
  for(size_t k=0;k<n;k++){ double q[3]; to3(loop[2*k], loop[2*k+1], q); sh.verts.insert(sh.verts.end(), q, q+3); }
  if( const Field *fp = find(s, "prism") ){
    const double d[3] = { num(*fp,0), num(*fp,1), num(*fp,2) };
    for(size_t k=0;k<n;k++) for(int i=0;i<3;i++) sh.verts.push_back(sh.verts[3*k+i] + d[i]);
    for(size_t k=0;k<n;k++){ const int a = (int)k, b = (int)((k+1)%n), a2 = a+(int)n, b2 = b+(int)n;
      tri.insert(tri.end(), {a, b, b2}); tri.insert(tri.end(), {a, b2, a2}); }           /* side quads */
    for(size_t k=1;k+1<n;k++){ tri.insert(tri.end(), {0, (int)k+1, (int)k}); tri.insert(tri.end(), {(int)n, (int)(n+k), (int)(n+k+1)}); }   /* caps as signed fans */
  } else if( const Field *fp = find(s, "pyramid") ){
    for(int i=0;i<3;i++) sh.verts.push_back(num(*fp,i));
    for(size_t k=0;k<n;k++) tri.insert(tri.end(), {(int)k, (int)((k+1)%n), (int)n});
    for(size_t k=1;k+1<n;k++) tri.insert(tri.end(), {0, (int)k+1, (int)k});
  } else { warn += "polyhedron loop without prism/pyramid; "; return true; }
  sh.mp = mesh_massprop(sh.verts, tri); sh.has_mp = true;
  return true;
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
    /* ax / ay / az: the box axes in the link frame (two given: the third is their cross product; one: completed) */
    { double ax[3][3] = {{1,0,0},{0,1,0},{0,0,1}}; bool have[3] = {false,false,false}; const char *key[3] = {"ax","ay","az"};
      for(int a=0;a<3;a++) if( const Field *f = find(s, key[a]) ){ double n = 0; for(int k=0;k<3;k++){ ax[a][k] = num(*f, k); n += ax[a][k]*ax[a][k]; }
        n = std::sqrt(n); if( n > 0 ){ for(int k=0;k<3;k++) ax[a][k] /= n; have[a] = true; } }
      auto cross3 = [](const double *u, const double *v, double *w){ w[0] = u[1]*v[2]-u[2]*v[1]; w[1] = u[2]*v[0]-u[0]*v[2]; w[2] = u[0]*v[1]-u[1]*v[0]; };
      if( have[0] || have[1] || have[2] ){
        if( have[0] && have[1] ) cross3(ax[0], ax[1], ax[2]);
        else if( have[1] && have[2] ) cross3(ax[1], ax[2], ax[0]);
        else if( have[2] && have[0] ) cross3(ax[2], ax[0], ax[1]);
        else { const int g = have[0] ? 0 : ( have[1] ? 1 : 2 ); double a_[3], e1[3], e2[3]; basis_perp(ax[g], a_, e1, e2);
          for(int k=0;k<3;k++){ ax[(g+1)%3][k] = e1[k]; ax[(g+2)%3][k] = e2[k]; } }
        for(int r=0;r<3;r++) for(int c=0;c<3;c++) b.R[3*r+c] = ax[c][r];
      } }
    sh.is_box = true; sh.box = b;
    for(int k=0;k<8;k++){ const double l[3] = { ((k&1)?0.5:-0.5)*b.depth, ((k&2)?0.5:-0.5)*b.width, ((k&4)?0.5:-0.5)*b.height };
      for(int r=0;r<3;r++) sh.verts.push_back(b.center[r] + b.R[3*r]*l[0] + b.R[3*r+1]*l[1] + b.R[3*r+2]*l[2]); }
    { const double V = b.depth*b.width*b.height, d2[3] = { b.depth*b.depth/12.0, b.width*b.width/12.0, b.height*b.height/12.0 };
      sh.mp.vol = V; for(int i=0;i<3;i++) sh.mp.vc[i] = V*b.center[i];
      for(int i=0;i<3;i++) for(int j=0;j<3;j++){ double t = 0; for(int k=0;k<3;k++) t += b.R[3*i+k]*d2[k]*b.R[3*j+k]; sh.mp.xx[3*i+j] = V*( t + b.center[i]*b.center[j] ); }
      sh.has_mp = true; }
    return true;
  }
  if( type == "polyhedron" ){
    if( find(s, "loop") ){ read_loop_sugar(s, sh, warn); return true; }
    std::vector<int> tri;
    for(const Field &f : s.fields){
      if( f.key == "vert" && f.val.size() >= 4 ){ sh.verts.push_back(num(f,1)); sh.verts.push_back(num(f,2)); sh.verts.push_back(num(f,3)); }
      if( f.key == "face" && f.val.size() >= 3 ) for(int k=0;k<3;k++) tri.push_back((int)num(f,k));
    }
    bool ok = !tri.empty(); for(int t : tri) if( t < 0 || 3*t+2 >= (int)sh.verts.size() ) ok = false;
    if( ok ){ sh.mp = mesh_massprop(sh.verts, tri); sh.has_mp = true; }
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
    double a[3], e1[3], e2[3];
    if( type == "cylinder" ){ const double ax[3] = {c[1][0]-c[0][0], c[1][1]-c[0][1], c[1][2]-c[0][2]}; ring(sh.verts, c[0], ax, r, div); ring(sh.verts, c[1], ax, r, div);
      const double h = std::sqrt(ax[0]*ax[0]+ax[1]*ax[1]+ax[2]*ax[2]), V = M_PI*r*r*h; basis_perp(ax, a, e1, e2);
      /* second moments about the centroid: transverse V r^2/4, axial V h^2/12 */
      sh.mp = axial_massprop(c[0], a, V, 0.5*h, V*r*r/4.0, V*h*h/12.0); sh.has_mp = true; }
    else { const double ax[3] = {vert[0]-c[0][0], vert[1]-c[0][1], vert[2]-c[0][2]}; if( hasv ){ sh.verts.push_back(vert[0]); sh.verts.push_back(vert[1]); sh.verts.push_back(vert[2]); } ring(sh.verts, c[0], ax, r, div);
      const double h = std::sqrt(ax[0]*ax[0]+ax[1]*ax[1]+ax[2]*ax[2]), V = M_PI*r*r*h/3.0; basis_perp(ax, a, e1, e2);
      /* cone: centroid h/4 above the base; second moments about it: transverse 3 V r^2/20, axial 3 V h^2/80 */
      sh.mp = axial_massprop(c[0], a, V, 0.25*h, 3.0*V*r*r/20.0, 3.0*V*h*h/80.0); sh.has_mp = true; }
    return true;
  }
  if( type == "sphere" ){
    double c[3] = {0,0,0}; if( const Field *f = find(s, "center") ) for(int k=0;k<3;k++) c[k] = num(*f, k);
    const Field *fr = find(s, "radius"); const double r = fr ? num(*fr, 0) : 0.0;
    /* [EXT] Zeo tessellates a sphere for the vertex test; its default division is not visible in the tree: 8 latitude
     * bands x 8 meridians here (58 vertices), `div` when the file gives one */
    const Field *fd = find(s, "div"); const int div = fd ? (int)num(*fd, 0) : 8;
    sh.verts.insert(sh.verts.end(), {c[0], c[1], c[2]+r});
    for(int i=1;i<div;i++){ const double ph = M_PI*i/div; for(int j=0;j<div;j++){ const double th = 2.0*M_PI*j/div;
      sh.verts.insert(sh.verts.end(), {c[0] + r*std::sin(ph)*std::cos(th), c[1] + r*std::sin(ph)*std::sin(th), c[2] + r*std::cos(ph)}); } }
    sh.verts.insert(sh.verts.end(), {c[0], c[1], c[2]-r});
    const double V = 4.0/3.0*M_PI*r*r*r; const double az[3] = {0,0,1};
    sh.mp = axial_massprop(c, az, V, 0.0, V*r*r/5.0, V*r*r/5.0); sh.has_mp = true;
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
      else if( jt == "breakablefloat" ) l.jtype = J_BRFLOAT;
      else if( jt == "cylindrical" ) l.jtype = J_CYLIN; else if( jt == "hooke" || jt == "universal" ) l.jtype = J_HOOKE;
      else { err = "joint type '" + jt + "' of link '" + l.name + "' is not supported"; return false; }
      if( const Field *f = find(s, "mass") ) l.mass = num(*f, 0);
      bool com_auto = false, inertia_auto = false; double density = 0.0;
      if( const Field *f = find(s, "density") ) density = num(*f, 0);
      if( const Field *f = find(s, "COM") ){ if( !f->val.empty() && f->val[0] == "auto" ) com_auto = true; else for(int k=0;k<3;k++) l.com[k] = num(*f, k); }
      if( const Field *f = find(s, "inertia") ){ if( !f->val.empty() && f->val[0] == "auto" ) inertia_auto = true; else for(int k=0;k<9;k++) l.inertia[k] = num(*f, k); }
      if( const Field *f = find(s, "frame") ) for(int r=0;r<3;r++){ for(int c=0;c<3;c++) l.Ro[3*r+c] = num(*f, 4*r+c, r==c); l.po[r] = num(*f, 4*r+3); }
      if( const Field *f = find(s, "pos") ) for(int k=0;k<3;k++) l.po[k] = num(*f, k);
      if( const Field *f = find(s, "DH") ){
        /* modified DH (a, alpha, d, theta), angles in degrees ([EXT] zFrame3DFromDH) */
        const double a = num(*f,0), al = num(*f,1)*DEG, d = num(*f,2), th = num(*f,3)*DEG;
        const double sa = std::sin(al), ca = std::cos(al), st = std::sin(th), ct = std::cos(th);
        const double R[9] = { ct, -st, 0,  ca*st, ca*ct, -sa,  sa*st, sa*ct, ca };
        std::memcpy(l.Ro, R, sizeof R); l.po[0] = a; l.po[1] = -d*sa; l.po[2] = d*ca;
      }
      if( const Field *f = find(s, "forcethreshold") ) l.brk_f = num(*f, 0);
      if( const Field *f = find(s, "torquethreshold") ) l.brk_t = num(*f, 0);
      if( const Field *f = find(s, "break") ){ l.brk_f = num(*f, 0); l.brk_t = num(*f, 1); }
      if( const Field *f = find(s, "stiffness") ) l.stiffness = num(*f, 0);
      if( const Field *f = find(s, "viscosity") ) l.viscosity = num(*f, 0);
      if( const Field *f = find(s, "coulomb") ) l.coulomb = num(*f, 0);
      if( const Field *f = find(s, "staticfriction") ) l.sfriction = num(*f, 0);
      const std::string mn = word(s, "motor");
      if( !mn.empty() ){ auto it = motors.find(mn); if( it == motors.end() ){ err = "unknown motor '" + mn + "'"; return false; } l.motor = it->second; }
      const std::string pn = word(s, "parent");
      if( !pn.empty() ){ auto it = link_index.find(pn); if( it == link_index.end() ){ err = "parent '" + pn + "' of link '" + l.name + "' must be defined before it"; return false; } l.parent = it->second; }
      MassProp lmp;
      for(const Field &f : s.fields) if( f.key == "shape" && !f.val.empty() ){
        auto it = shapes.find(f.val[0]); if( it == shapes.end() ){ err = "unknown shape '" + f.val[0] + "'"; return false; }
        if( !it->second.verts.empty() ) l.shapes.push_back(it->second.verts);
        if( it->second.is_box ){ BoxShape bx = it->second.box; bx.cloud = it->second.verts.empty() ? -1 : (int)l.shapes.size() - 1; l.boxes.push_back(bx); }
        if( it->second.has_mp ) lmp.add(it->second.mp); else if( com_auto || inertia_auto ) warn += "shape '" + f.val[0] + "' has no mass properties for COM/inertia: auto; ";
      }
      /* `COM: auto` / `inertia: auto` / `density:`: uniform density over the link's shapes ([EXT] Zeo; arm.ztk:79-80) */
      if( ( com_auto || inertia_auto || density > 0.0 ) && lmp.vol > 0.0 ){
        if( density > 0.0 && !( l.mass > 0.0 ) ) l.mass = density*lmp.vol;
        const double rho = l.mass/lmp.vol; double c[3]; for(int k=0;k<3;k++) c[k] = lmp.vc[k]/lmp.vol;
        if( com_auto || find(s, "COM") == nullptr ) for(int k=0;k<3;k++) l.com[k] = c[k];
        if( inertia_auto || find(s, "inertia") == nullptr ){
          /* inertia about the COM from the second moments about the origin: I = tr(X) E - X with X = rho xx - m c c^T */
          double X[9]; for(int i=0;i<3;i++) for(int j=0;j<3;j++) X[3*i+j] = rho*lmp.xx[3*i+j] - l.mass*l.com[i]*l.com[j];
          const double tr = X[0]+X[4]+X[8];
          for(int i=0;i<3;i++) for(int j=0;j<3;j++) l.inertia[3*i+j] = (i==j ? tr : 0.0) - X[3*i+j];
        }
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
        const bool angular = l.jtype == J_REVOL || l.jtype == J_SPHER || l.jtype == J_HOOKE || ( ( l.jtype == J_FLOAT || l.jtype == J_BRFLOAT ) && k >= 3 ) || ( l.jtype == J_CYLIN && k == 1 );
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
