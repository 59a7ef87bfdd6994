/* Host check of the exact 32-bit modulo the kernels use (bsg_device.cuh: mod_m32; proof in DESIGN.md §5):
 * the same six operations in plain C against 128-bit arithmetic, over edge moduli (1, 2, 2^k, 2^k - 1, 2^30 - 1,
 * random widths) and edge values (0, m-1, m, multiples of m +- 1, 2^64 - 1, all-ones halves).
 * Built and run by tests/test_host_logic.py::test_mod_m32_formula (gcc, no GPU). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
static inline uint32_t umin(uint32_t a, uint32_t b){return a<b?a:b;}
static uint32_t mod_m32(uint64_t x, uint32_t m, uint32_t ih, uint32_t il){
    uint32_t xh=x>>32, xl=(uint32_t)x, nm=0u-m;
    uint64_t s=(uint64_t)xl*ih; s+= (uint64_t)xh*il; // wraps mod 2^64
    uint32_t q=(uint32_t)(s>>32)+xh*ih;
    uint32_t r=q*nm+xl;
    r=umin(r,r-m); r=umin(r,r-m); return r;
}
static uint64_t rnd_state=88172645463325252ull;
static uint64_t rnd(){rnd_state^=rnd_state<<13;rnd_state^=rnd_state>>7;rnd_state^=rnd_state<<17;return rnd_state;}
int main(){
    uint64_t bad=0, n=0;
    for(int it=0; it<6000; ++it){
        uint32_t m;
        if(it<64) m=it+1;
        else if(it<64+60){int k=(it-64)/2+1; m=(it&1)?(1u<<k)-1:(1u<<k); if(m==0)m=1; if(m>=(1u<<30)) m=(1u<<30)-1;}
        else if(it<200) m=(1u<<30)-1-(it-124);
        else { int bits=1+rnd()%30; m=(uint32_t)(rnd()&((1u<<bits)-1)); if(m==0)m=1; if(m>=(1u<<30)) m=(1u<<30)-1; }
        unsigned __int128 one=1; uint64_t I = m==1? ~0ull : (uint64_t)((one<<64)/m);
        uint32_t ih=I>>32, il=(uint32_t)I;
        for(int j=0;j<4000;++j){
            uint64_t x;
            switch(j){case 0:x=0;break;case 1:x=m-1;break;case 2:x=m;break;case 3:x=~0ull;break;case 4:x=(uint64_t)m<<32;break;case 5:x=((uint64_t)m<<32)-1;break;
              case 6:x=0xffffffffull;break;case 7:x=0xffffffff00000000ull;break;case 8:x=~0ull-m;break;case 9:x=(~0ull/m)*m;break;case 10:x=(~0ull/m)*m-1;break;
              default: x=rnd(); if(j&1) x|=0xffffffff00000000ull; if((j&6)==6) x|=0xffffffffull; }
            uint32_t r=mod_m32(x,m,ih,il);
            if(r!=(uint32_t)(x%m)){ if(bad<10) printf("BAD m=%u x=%llu got %u want %llu\n",m,(unsigned long long)x,r,(unsigned long long)(x%m)); ++bad;}
            ++n;
        }
    }
    printf("%llu checks, %llu bad\n",(unsigned long long)n,(unsigned long long)bad);
    return bad!=0;
}
