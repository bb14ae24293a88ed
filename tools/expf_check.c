#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static const uint64_t T[32]={0x3ff0000000000000,0x3fefd9b0d3158574,0x3fefb5586cf9890f,0x3fef9301d0125b51,0x3fef72b83c7d517b,0x3fef54873168b9aa,0x3fef387a6e756238,0x3fef1e9df51fdee1,0x3fef06fe0a31b715,0x3feef1a7373aa9cb,0x3feedea64c123422,0x3feece086061892d,0x3feebfdad5362a27,0x3feeb42b569d4f82,0x3feeab07dd485429,0x3feea47eb03a5585,0x3feea09e667f3bcd,0x3fee9f75e8ec5f74,0x3feea11473eb0187,0x3feea589994cce13,0x3feeace5422aa0db,0x3feeb737b0cdc5e5,0x3feec49182a3f090,0x3feed503b23e255d,0x3feee89f995ad3ad,0x3feeff76f2fb5e47,0x3fef199bdd85529c,0x3fef3720dcef9069,0x3fef5818dcfba487,0x3fef7c97337b9b5f,0x3fefa4afa2a490da,0x3fefd0765b6e4540};
static inline double asd(uint64_t u){double d;memcpy(&d,&u,8);return d;}
static inline uint64_t asu(double d){uint64_t u;memcpy(&u,&d,8);return u;}
#define N 32
static const double InvLn2N = 0x1.71547652b82fep+0 * N;
static const double SHIFT = 0x1.8p+52;
static const double C0 = 0x1.c6af84b912394p-5/N/N/N, C1 = 0x1.ebfce50fac4f3p-3/N/N, C2 = 0x1.62e42ff0c52d6p-1/N;
float my_expf(float x, int use_fma, int roundint){
  double xd=x, z=InvLn2N*xd, kd; uint64_t ki;
  if (roundint){ kd = rint(z); ki=(uint64_t)(int64_t)kd; } else { kd = z+SHIFT; ki=asu(kd); kd-=SHIFT; }
  double r=z-kd;
  uint64_t t=T[ki%N]; t+=ki<<(52-5); double s=asd(t);
  double zz,r2,y;
  if(use_fma){ zz=fma(C0,r,C1); r2=r*r; y=fma(C2,r,1.0); y=fma(zz,r2,y); y=y*s; }
  else { volatile double a=C0*r; zz=a+C1; r2=r*r; volatile double b=C2*r; y=b+1.0; volatile double c=zz*r2; y=c+y; y=y*s; }
  return (float)y;
}
int main(){
  long n=20000000, bad_f=0,bad_n=0,bad_fr=0; srand(1);
  for(long i=0;i<n;i++){
    uint32_t u=((uint32_t)rand()<<16)^rand();
    float x=-2.2f*(float)((double)u/4294967296.0);
    float ref=expf(x);
    if(my_expf(x,1,0)!=ref) bad_f++;
    if(my_expf(x,0,0)!=ref) bad_n++;
    if(my_expf(x,1,1)!=ref) bad_fr++;
  }
  printf("n=%ld mismatches: fma(shift)=%ld nofma(shift)=%ld fma(rint)=%ld\n",n,bad_f,bad_n,bad_fr);
  // correctly-rounded via double exp
  long bad_d=0; srand(1);
  for(long i=0;i<n;i++){ uint32_t u=((uint32_t)rand()<<16)^rand(); float x=-2.2f*(float)((double)u/4294967296.0); if((float)exp((double)x)!=expf(x)) bad_d++; }
  printf("double-exp-rounded mismatches vs expf: %ld\n",bad_d);
  return 0;}
