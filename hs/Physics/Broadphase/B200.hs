{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE RecordWildCards          #-}

{- |
Binding of libshapes_b200.so (include/shapes_b200.h) for ublubu/shapes.

NOT COMPILED in this repository: GHC is not part of the build environment.  It is the source a
maintainer of the reference would add next to Physics.Broadphase.Aabb / Grid; the same C ABI is
exercised by the Python ctypes binding (shapes_b200/_lib.py) and the C++ mirror
(include/shapes_b200.hpp) in tests/.  See INTEGRATION.md for where it plugs into
Physics.Engine.Main.updateWorld (shapes/src/Physics/Engine/Main.hs:71-86).

Marshalling notes
  * _wPhysObjs :: U.MVector s PhysicalObj unboxes to one Double column per field
    (Physics/Constraint.hs:52-63, Physics/Linear.hs:49-52): the columns are copied into pinned
    Storable buffers (shapes_host_alloc) once per frame; hull geometry (a boxed vector of boxed
    arrays, Physics/World.hs:50) is flattened once into CSR columns by 'setShapes'.
  * cos / sin of the rotation are computed on the host with the same libm calls rotate22 makes
    (Physics/Linear.hs:353-357) and passed down, which keeps every result bit-identical.
  * The engine monad is ST: calls are wrapped with unsafeIOToST.  The frame call is `safe`
    (blocking, other Haskell threads keep running); everything else is `unsafe`.
-}
module Physics.Broadphase.B200
  ( Ctx, Frame(..), StepConfig(..)
  , create, destroy, setShapes, frame, growAndRetry, c_grow
  , Multi, c_createMulti, c_multiDestroy, c_multiSetShapes, c_multiFrame
  , worldUpload, worldStep, worldDownload, sincos
  ) where

import           Control.Monad.ST
import           Control.Monad.ST.Unsafe      (unsafeIOToST)
import           Data.Int
import qualified Data.Vector.Storable         as S
import qualified Data.Vector.Unboxed          as V
import           Data.Word
import           Foreign
import           Foreign.C.String
import           Foreign.C.Types

import           Physics.Constraint           (PhysicalObj (..))
import           Physics.Constraints.Contact  (ObjectFeatureKey (..))
import           Physics.Constraints.Types    (ContactConstraint (..), RestitutionConstraint (..))
import           Physics.Contact.Types        (Contact (..), ContactBehavior (..))
import           Physics.World                (World)
import           Utils.Descending             (Descending (..))
import           Utils.Utils                  (Flipping (..))

data Ctx          -- opaque shapes_ctx
data FrameOut     -- shapes_frame_out; field offsets through hsc2hs (#peek / #poke) in a real build
data StepStats    -- shapes_step_stats

-- | shapes_step_config: EngineConfig + ContactBehavior + the External applied each frame.
data StepConfig = StepConfig
  { scDt, scBaumgarte, scSlop :: !Double
  , scExternalKind            :: !Int32   -- 0 none, 1 constantAccel, 2 constantForce (World/External.hs:16-28)
  , scSolverIterations        :: !Int32   -- improveWorld sweeps; updateWorld runs 2
  , scExternalX, scExternalY  :: !Double
  , scWarmStart               :: !Int32
  }

-- | What one shapes_frame call returns, in the reference's own types.
data Frame = Frame
  { frameKeys        :: Descending (Int, Int)                                  -- G.culledKeys / Aabb.culledKeys
  , frameContacts    :: Descending (ObjectFeatureKey Int, Flipping Contact)    -- prepareFrame
  , frameConstraints :: V.Vector ContactConstraint                             -- constraintGen, row k <-> contact k
  }

foreign import ccall unsafe "shapes_create"
  c_create :: Ptr (Ptr Ctx) -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> IO CInt
foreign import ccall unsafe "shapes_destroy"
  c_destroy :: Ptr Ctx -> IO ()
foreign import ccall unsafe "shapes_last_error"
  c_lastError :: Ptr Ctx -> IO CString
foreign import ccall unsafe "shapes_set_shapes"
  c_setShapes :: Ptr Ctx -> Int64 -> Ptr Word8 -> Ptr Int32 -> Ptr Double -> Ptr Double
              -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> IO CInt
foreign import ccall safe "shapes_frame"
  c_frame :: Ptr Ctx -> Int64
          -> Ptr Double -> Ptr Double                 -- pos_x pos_y   (_physObjPos)
          -> Ptr Double -> Ptr Double -> Ptr Double   -- rot, cos rot, sin rot
          -> Ptr Double -> Ptr Double                 -- inv_lin inv_rot (_physObjInvMass)
          -> Double -> Double -> Double               -- dt, contactBaumgarte, contactPenetrationSlop
          -> Ptr FrameOut -> IO CInt
foreign import ccall unsafe "shapes_set_lagrangian_cache"
  c_setCache :: Ptr Ctx -> Int64 -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall unsafe "shapes_host_alloc" c_hostAlloc :: CSize -> IO (Ptr a)
foreign import ccall unsafe "shapes_host_free"  c_hostFree  :: Ptr a -> IO ()

-- device-resident world (optional: the whole updateWorld on the GPU)
foreign import ccall safe "shapes_world_upload"
  c_worldUpload :: Ptr Ctx -> Int64
                -> Ptr Double -> Ptr Double -> Ptr Double      -- _physObjVel x/y, _physObjRotVel
                -> Ptr Double -> Ptr Double -> Ptr Double      -- _physObjPos x/y, _physObjRotPos
                -> Ptr Double -> Ptr Double                    -- cos/sin the shapes were last moved with (nullPtr: shapes_sincos)
                -> Ptr Double -> Ptr Double                    -- _imLin, _imRot
                -> Ptr Double -> Ptr Double -> IO CInt         -- _mMu, _mBounce
foreign import ccall safe "shapes_world_step"
  c_worldStep :: Ptr Ctx -> Ptr StepConfig -> Ptr StepStats -> IO CInt
foreign import ccall safe "shapes_world_download"
  c_worldDownload :: Ptr Ctx -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double
                  -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall unsafe "shapes_sincos"
  c_sincos :: Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> IO ()

-- | Every non-zero code except SHAPES_E_CAPACITY becomes 'error': the replaced functions are total
-- and have no error channel (Physics/Solvers/Contact.hs:40-52).
check :: Ptr Ctx -> CInt -> IO ()
check _   0  = return ()
check ctx rc = c_lastError ctx >>= peekCString >>= \msg -> error ("shapes_b200 " ++ show rc ++ ": " ++ msg)

create :: Int -> Int -> Int -> Int -> IO (Ptr Ctx)
create maxShapes maxVerts maxPairs maxContacts = alloca $ \out -> do
  rc <- c_create out 0 (fromIntegral maxShapes) (fromIntegral maxVerts) (fromIntegral maxPairs) (fromIntegral maxContacts)
  if rc /= 0 then c_lastError nullPtr >>= peekCString >>= error else peek out

destroy :: Ptr Ctx -> IO ()
destroy = c_destroy

-- | World.fromList / append / delete (Physics/World.hs:77-116): filled flags of the EmptiesVector,
-- CSR offsets, local CCW vertices (_hullLocalVertices) and, for CircleShapes, the radius column.
setShapes :: Ptr Ctx -> S.Vector Word8 -> S.Vector Int32 -> S.Vector Double -> S.Vector Double -> Maybe (S.Vector Double) -> IO ()
setShapes ctx alive offs lx ly radius =
  S.unsafeWith alive $ \pa -> S.unsafeWith offs $ \po -> S.unsafeWith lx $ \px -> S.unsafeWith ly $ \py ->
    maybe ($ nullPtr) S.unsafeWith radius $ \pr ->
      c_setShapes ctx (fromIntegral (S.length alive)) pa po px py nullPtr nullPtr pr >>= check ctx

-- | One frame: the three expressions of updateWorld this library replaces.
frame :: Ptr Ctx -> ContactBehavior -> Double -> World s label -> ST s Frame
frame ctx ContactBehavior{..} dt world = unsafeIOToST $ do
  cols <- marshalBodies world                  -- pos, rot, cos, sin, inverse masses (pinned)
  withFrameOut $ \out -> do
    rc <- c_frame ctx (nSlots cols) (posX cols) (posY cols) (rot cols) (cosRot cols) (sinRot cols)
                  (invLin cols) (invRot cols) dt contactBaumgarte contactPenetrationSlop out
    case rc of
      0  -> readFrame out
      -4 -> growAndRetry ctx out               -- SHAPES_E_CAPACITY: the required sizes are in `out`
      _  -> check ctx rc >> undefined

-- | SHAPES_E_CAPACITY: the failed call left the ctx as it was (previous frame's keys, the Lagrangian cache, an
-- uploaded world); shapes_grow enlarges the capacities in place and the same call is issued again, so the retried
-- frame still joins against the EngineCache applyCachedSlns would have used (Physics/Solvers/Contact.hs:84-121).
foreign import ccall unsafe "shapes_grow"
  c_grow :: Ptr Ctx -> Int64 -> Int64 -> IO CInt

-- | One process, several GPUs (the host is one single-threaded ST computation, Physics/Engine/Main.hs:38,71-86):
-- same arguments as shapes_frame, `out` receives the whole frame in the reference's descending order.
data Multi
foreign import ccall unsafe "shapes_create_multi"
  c_createMulti :: Ptr (Ptr Multi) -> CInt -> Ptr CInt -> Int64 -> Int64 -> Int64 -> Int64 -> IO CInt
foreign import ccall unsafe "shapes_multi_destroy"
  c_multiDestroy :: Ptr Multi -> IO ()
foreign import ccall safe "shapes_multi_set_shapes"
  c_multiSetShapes :: Ptr Multi -> Int64 -> Ptr Word8 -> Ptr Int32 -> Ptr Double -> Ptr Double
                   -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> IO CInt
foreign import ccall safe "shapes_multi_frame"
  c_multiFrame :: Ptr Multi -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double
               -> Ptr Double -> Ptr Double -> Double -> Double -> Double -> Ptr FrameOut -> IO CInt

-- Compact wire format: withFrameOut leaves the pointers of the sixteen derived constraint columns NULL (nothing is
-- copied for a NULL column) and readFrame rebuilds them, bit for bit, from what is shipped
--   key_i key_j feat_a feat_b flip | normal_x/y center_x/y depth | j_np2 j_np5 j_f2 j_f5 | b_np | inv_eff_np inv_eff_f:
-- with n = normal, s = if flip then n else negateV2 n, c = center
--   _ccNonPen      = Constraint (V6 s.x s.y j_np2 (-s.x) (-s.y) j_np5) b_np      (NonPenetration.hs:34-43; flip3v3 for Flip)
--   _ccFriction    = Constraint (V6 s.y (-s.x) j_f2 (-s.y) s.x j_f5) 0           (Friction.hs:31-44)
--   _ccRestitution = RestitutionConstraint (c - pos i) (c - pos j) (negateV2 s)  (Restitution.hs:21-31)
-- 113 B per contact row cross PCIe instead of 225.
--
-- marshalBodies / withFrameOut / readFrame / growAndRetry: buffer management and the hsc2hs
-- peeks of shapes_frame_out; readFrame builds
--   Descending [(i, j)]                                                 from pair_i / pair_j
--   Descending [(ObjectFeatureKey (i, j) (fa, fb), Same c | Flip c)]    from the contact columns (flip == 0 -> Same)
--   V.Vector ContactConstraint                                          from the constraint columns:
--       _ccNonPen = Constraint (V6 j_np0..5) b_np, _ccRestitution = RestitutionConstraint (V2 ra) (V2 rb) (V2 rn),
--       _ccFriction = Constraint (V6 j_f0..5) 0
-- (Physics/Constraints/Types.hs:41-51).  They are omitted here because they cannot be type-checked
-- without GHC; the Python and C++ mirrors in this repository implement exactly that unpacking.
