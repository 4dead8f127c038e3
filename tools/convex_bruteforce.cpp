// Brute-force validation of include/b200mj_convex.h: for random pairs of convex primitives, the penetration depth of cvx_pair
// against the minimum over 80 000 random + polished directions of the overlap function h(d). Development tool:
//   g++ -O2 -o build/convex_bruteforce tools/convex_bruteforce.cpp && build/convex_bruteforce
#include <stdio.h>
#include <stdlib.h>
#include "../include/b200mj_convex.h"   // (path relative to tools/)
static void rotm(double* R) { // random rotation
  double q[4]; double n=0; for(int i=0;i<4;i++){q[i]=drand48()*2-1;n+=q[i]*q[i];} n=sqrt(n); for(int i=0;i<4;i++)q[i]/=n;
  double w=q[0],x=q[1],y=q[2],z=q[3];
  R[0]=1-2*(y*y+z*z);R[1]=2*(x*y-w*z);R[2]=2*(x*z+w*y);R[3]=2*(x*y+w*z);R[4]=1-2*(x*x+z*z);R[5]=2*(y*z-w*x);R[6]=2*(x*z-w*y);R[7]=2*(y*z+w*x);R[8]=1-2*(x*x+y*y);
}
int main() {
  srand48(3);
  int types[5]={2,3,4,5,6}; double sizes[5][3]={{.2,0,0},{.1,.2,0},{.3,.2,.1},{.15,.25,0},{.2,.15,.1}};
  int nbad=0, n=0;
  for (int trial=0; trial<3000; trial++) {
    int ia=lrand48()%5, ib=lrand48()%5; if (ia>ib){int t=ia;ia=ib;ib=t;}
    if (types[ib]<=3) continue; if (types[ia]==3&&types[ib]==6) continue;
    double Ra[9],Rb[9]; rotm(Ra); rotm(Rb);
    double pa[3]={0,0,0}, pb[3]={drand48()-.5,drand48()-.5,drand48()-.5};
    double dist,pos[3],nr[3];
    int r=cvx_pair(types[ia],pa,Ra,sizes[ia],types[ib],pb,Rb,sizes[ib],0,&dist,pos,nr);
    CvxGeom g1={types[ia],pa,Ra,sizes[ia],0}, g2={types[ib],pb,Rb,sizes[ib],0};
    double best=1e9, bd[3];
    for (int i=0;i<60000;i++){ double d[3]={drand48()*2-1,drand48()*2-1,drand48()*2-1}; cvx_normalize(d); CvxSup s; cvx_support(g1,g2,d,s); double h=cvx_dot(s.v,d); if(h<best){best=h;bd[0]=d[0];bd[1]=d[1];bd[2]=d[2];} }
    // local polish of brute force
    for (int k=0;k<20000;k++){ double d[3]={bd[0]+(drand48()-.5)*.02,bd[1]+(drand48()-.5)*.02,bd[2]+(drand48()-.5)*.02}; cvx_normalize(d); CvxSup s; cvx_support(g1,g2,d,s); double h=cvx_dot(s.v,d); if(h<best){best=h;bd[0]=d[0];bd[1]=d[1];bd[2]=d[2];} }
    if ((best>0) != (r==1)) { if (fabs(best)>1e-5) {printf("MISS types %d %d best %g r %d dist %g\n", types[ia],types[ib],best,r,r?dist:0.0); nbad++;} continue; }
    if (!r) continue;
    if (best>0.05) continue;
    n++;
    double err=(-dist)-best; double nd=1-(nr[0]*bd[0]+nr[1]*bd[1]+nr[2]*bd[2]);
    if (err>2e-4+0.02*best || err<-1e-4) { printf("DEPTH types %d %d true %g got %g ndot %g\n", types[ia],types[ib],best,-dist,nd); nbad++; }
  }
  printf("checked %d shallow hits, bad %d\n", n, nbad);
}
