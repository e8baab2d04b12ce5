/* boxdrop.c - a plain C caller of the rkfd_sim step API, written the way the reference's example programs use it
 * (reference example/chain/boxdrop_test.c: create, register chains from ZTK files, scan contact info, choose integrator,
 * properties and solver through the header's macros, UpdateInit, Update loop reading joint values through the chain,
 * UpdateDestroy, Destroy).  It includes ONE header and links against librokifd_b200.so: no CUDA in sight.
 *
 *   boxdrop <cube.ztk> <floor.ztk> <contacts.ztk|-> <Vert|MLCP|Volume> <steps> <z0>
 * prints the six joint displacements of the cube after every 10th step and at the end ("final: ..."). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <roki_fd/rkfd_b200.h>

int main(int argc, char *argv[])
{
  rkFD fd;
  rkFDCell *cube;
  zVec dis, vel;
  int i, step, nstep;

  if( argc < 7 ){ fprintf( stderr, "usage: %s cube.ztk floor.ztk contacts.ztk|- Vert|MLCP|Volume steps z0\n", argv[0] ); return 2; }
  nstep = atoi( argv[5] );
  rkFDCreate( &fd );
  if( !( cube = rkFDChainRegFile( &fd, argv[1] ) ) ) return 1;
  if( !rkFDChainRegFile( &fd, argv[2] ) ) return 1;
  if( strcmp( argv[3], "-" ) != 0 && !rkFDContactInfoScanFile( &fd, argv[3] ) ) return 1;
  rkFDODE2Assign( &fd, Regular );
  rkFDODE2AssignRegular( &fd, RKG );
  rkFDPrpSetDT( &fd, 0.001 );
  if( strcmp( argv[4], "MLCP" ) == 0 ) rkFDSetSolver( &fd, MLCP );
  else if( strcmp( argv[4], "Volume" ) == 0 ) rkFDSetSolver( &fd, Volume );
  else rkFDSetSolver( &fd, Vert );

  dis = zVecAlloc( rkChainJointSize( rkFDCellChain(cube) ) );
  vel = zVecAlloc( rkChainJointSize( rkFDCellChain(cube) ) );
  zVecZero( dis ); zVecZero( vel );
  zVecSetElem( dis, 2, atof( argv[6] ) );
  zVecSetElem( dis, 3, 0.05 ); zVecSetElem( dis, 4, 0.02 );
  zVecSetElem( vel, 0, 0.3 );
  rkFDChainSetDis( cube, dis );
  rkFDChainSetVel( cube, vel );

  rkFDUpdateInit( &fd );
  for( step=1; step<=nstep; step++ ){
    rkFDUpdate( &fd );
    if( step % 10 == 0 || step == nstep ){
      rkChainGetJointDisAll( rkFDCellChain(cube), dis );
      printf( step == nstep ? "final:" : "%d:", step );
      for( i=0; i<zVecSizeNC(dis); i++ ) printf( " %.17g", zVecElemNC(dis,i) );
      printf( " t=%.17g\n", rkFDTime(&fd) );
    }
  }
  rkFDUpdateDestroy( &fd );
  rkFDDestroy( &fd );
  zVecFree( dis ); zVecFree( vel );
  return 0;
}
