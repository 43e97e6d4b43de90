{-# LANGUAGE BangPatterns #-}
-- | Drop-in for "Data.Text.AhoCorasick.Automaton" (reference: src/Data/Text/AhoCorasick/Automaton.hs).
-- Same export list (:32-44) minus the debug helpers (`debugBuildDot`, `needleCasings`); the state-transition loop
-- (:442-534) runs in libam_b200.so on a B200.  NOT COMPILED HERE (no GHC in the image) -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Automaton
  ( AcMachine (..), CaseSensitivity (..), CodeUnitIndex (..), Match (..), Next (..)
  , build, runText, runLower, runWithCase
  ) where

import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (CodeUnitIndex (..), Text)
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal (alloca, allocaArray, peekArray)
import Foreign.Ptr (nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Vector as Vector

import Data.Text.AhoCorasick.FFI

data Match v = Match { matchPos :: {-# UNPACK #-} !CodeUnitIndex, matchValue :: v }   -- Automaton.hs:98-105
data Next a = Done !a | Step !a                                                        -- :398

-- | The native handle (needles + the host's toLower table; one device image per case mode, built at first use)
-- plus the host-side payload table: the native side reports needle indices, `machineValues` maps them back to
-- the caller's `v`.  Like the reference's `AcMachine` it knows nothing of case: `runText` and `runLower` run the
-- same machine (:539-553).
data AcMachine v = AcMachine
  { machineHandle :: !(ForeignPtr AmAutomaton)
  , machineValues :: !(Vector.Vector v)
  }

instance Functor AcMachine where                     -- the reference derives Functor (:123)
  fmap f m = m { machineValues = Vector.map f (machineValues m) }

-- | `build :: [(Text, v)] -> AcMachine v` (:176).
build :: [(Text, v)] -> AcMachine v
build needlesWithValues = unsafePerformIO $
  withSlices (map fst needlesWithValues) $ \needles n ->
  withLowerTable $ \lowerPtr ->
  alloca $ \out -> do
    _ <- amCall "am_automaton_build" [] $ c_am_automaton_build needles n lowerPtr nullPtr out
    h <- peek out >>= newForeignPtr c_am_automaton_free_ptr
    pure $ AcMachine h (Vector.fromList (map snd needlesWithValues))
{-# NOINLINE build #-}

-- | `runWithCase` (:443): left fold over all matches in the reference's order, honouring `Done`: the device returns
-- the ordered match array, the fold consumes it lazily and simply stops.
runWithCase :: CaseSensitivity -> a -> (a -> Match v -> Next a) -> AcMachine v -> Text -> a
runWithCase cs seed f machine text = go seed (findAll cs machine text)
  where
    go !acc [] = acc
    go !acc (AmMatch pos i : ms) =
      case f acc (Match (CodeUnitIndex (fromIntegral pos)) (machineValues machine `Vector.unsafeIndex` fromIntegral i)) of
        Step acc' -> go acc' ms
        Done r    -> r

runText, runLower :: a -> (a -> Match v -> Next a) -> AcMachine v -> Text -> a
runText  = runWithCase CaseSensitive   -- :539-541
runLower = runWithCase IgnoreCase      -- :551-553: the caller has lower-cased the needles

-- | am_find_all with the overflow protocol: on AM_E_OVERFLOW the call reports the capacity it needs; retry once.
findAll :: CaseSensitivity -> AcMachine v -> Text -> [AmMatch]
findAll cs machine text = unsafePerformIO $
  withForeignPtr (machineHandle machine) $ \h -> withSlice text $ \hay -> alloca $ \nFound ->
    let attempt cap = allocaArray cap $ \buf -> do
          rc <- amCall "am_find_all" [amEOverflow] $ c_am_find_all h (caseToC cs) hay buf (fromIntegral cap) nFound
          n <- fromIntegral <$> peek nFound
          if rc == amEOverflow then attempt n else peekArray n buf
    in attempt 4096
